#!/usr/bin/env python
"""Turn ncu outputs under gpurun_out/ into the small text summaries committed under profiles/.

  python scripts/ncu_summaries.py launches gpurun_out/launches.csv  profiles/launches_rNN.txt
  python scripts/ncu_summaries.py full     gpurun_out/x.ncu-rep     profiles/x_rNN.txt
"""
import collections
import csv
import re
import subprocess
import sys


def launches(src, dst):
    with open(src) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
    per = collections.OrderedDict()
    seq = []
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        t = float(r["Metric Value"]) / 1e3
        seq.append((int(r["ID"]), name, t, r["Grid Size"], r["Block Size"]))
        d = per.setdefault(name, [0, 0.0, 1e30, 0.0])
        d[0] += 1; d[1] += t; d[2] = min(d[2], t); d[3] = max(d[3], t)
    total = sum(d[1] for d in per.values())
    with open(dst, "w") as o:
        o.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src}); times are cold-cache and serialised\n")
        o.write(f"# {len(seq)} launches, total {total:.1f} us\n")
        o.write(f"{'kernel':62s} {'n':>5s} {'sum_us':>10s} {'share':>7s} {'min_us':>8s} {'max_us':>8s}\n")
        for name, d in sorted(per.items(), key=lambda kv: -kv[1][1]):
            o.write(f"{name[:62]:62s} {d[0]:5d} {d[1]:10.1f} {100 * d[1] / total:6.1f}% {d[2]:8.2f} {d[3]:8.2f}\n")
        # the last complete training step: from after the previous reduce_adam/adam to the last one
        ends = [i for i, s in enumerate(seq) if "adam" in s[1]]
        if len(ends) >= 2:
            o.write("\n# last training step in the capture (launch by launch)\n")
            st = 0.0
            for i in range(ends[-2] + 1, ends[-1] + 1):
                o.write(f"{seq[i][2]:9.2f} us  {seq[i][1][:70]:70s} grid={seq[i][3]} block={seq[i][4]}\n")
                st += seq[i][2]
            o.write(f"# step total {st:.1f} us\n")


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "sm__cycles_elapsed.max"]


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    hdr, units = r[0], r[1]
    with open(dst, "w") as o:
        o.write(f"# ncu --set full --clock-control none  ({src}), one block per captured launch\n")
        for row in r[2:]:
            o.write(f"\nkernel: {row[hdr.index('Kernel Name')]}\n")
            for k in METRICS:
                if k in hdr:
                    i = hdr.index(k)
                    o.write(f"  {k:72s} {row[i]:>16s} {units[i]}\n")
        srcp = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"],
                              capture_output=True, text=True).stdout
        rows = list(csv.reader(srcp.splitlines()))
        if len(rows) > 2:
            h = rows[1]
            blk = []
            for rr in rows[2:]:
                if rr and rr[0] == "Kernel Name":
                    break
                blk.append(rr)
            stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = collections.Counter()
            mix = collections.Counter()
            for rr in blk:
                try:
                    e = int(rr[h.index("Instructions Executed")])
                except Exception:
                    continue
                for s in stalls:
                    try:
                        tot[s] += int(rr[h.index(s)])
                    except Exception:
                        pass
                toks = rr[h.index("Source")].split()
                op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "?")
                mix[op.split(".")[0]] += e
            o.write("\n# first launch: warp-stall samples by reason\n")
            for s, v in tot.most_common(8):
                o.write(f"  {s:28s} {v}\n")
            te = sum(mix.values())
            o.write(f"\n# first launch: executed warp-instructions by opcode (total {te})\n")
            for k, v in mix.most_common(12):
                o.write(f"  {k:12s} {v:10d} {100 * v / max(te, 1):5.1f}%\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
