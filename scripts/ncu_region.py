#!/usr/bin/env python
"""Warp-stall samples of the tensor-core backward region of the fused fp32 kernel, from an `ncu --set full --import-source on`
capture (SASS page): the instructions between the first and the last HMMA, by stall reason and by opcode.

    python scripts/ncu_region.py gpurun_out/fused_full_s4.ncu-rep profiles/fused_mma_region_r02.txt"""
import collections
import csv
import subprocess
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], rows[2:]
    i_src, i_smp, i_exe = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    reasons = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
    hm = [i for i, r in enumerate(data) if "HMMA" in r[i_src]]
    lo, hi = hm[0] - 400, hm[-1] + 200
    region = data[lo:hi]

    def count(sel):
        c = collections.Counter()
        for r in sel:
            for i in reasons:
                try:
                    c[hdr[i]] += int(r[i])
                except ValueError:
                    pass
        return c

    total = sum(int(r[i_smp]) for r in data)
    reg = sum(int(r[i_smp]) for r in region)
    with open(dst, "w") as o:
        o.write(f"# {src}: {rows[0][1][:90]}\n")
        o.write(f"# {len(data)} SASS instructions, {total} warp-stall samples; region = instructions {lo}..{hi} (first HMMA - 400 .. last HMMA + 200)\n")
        o.write(f"region samples {reg} = {100 * reg / total:.1f} % of the kernel's; static HMMA instructions {len(hm)}, executed "
                f"{sum(int(data[i][i_exe]) for i in hm)}\n\n")
        for title, sel in (("HMMA instructions", [data[i] for i in hm]),
                           ("other instructions of the region", [r for r in region if "HMMA" not in r[i_src]]),
                           ("rest of the kernel", data[:lo] + data[hi:])):
            c = count(sel)
            n = sum(c.values())
            o.write(f"{title}: {n} samples\n")
            for k, v in c.most_common(8):
                o.write(f"  {k:28s} {v:6d}  {100 * v / max(n, 1):5.1f} %\n")
            o.write("\n")
        ops_s, ops_e = collections.Counter(), collections.Counter()
        for r in region:
            t = r[i_src].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops_s[op] += int(r[i_smp]); ops_e[op] += int(r[i_exe])
        o.write("region by opcode: samples, executed warp instructions\n")
        for k, v in ops_s.most_common(12):
            o.write(f"  {k:10s} {v:6d} {ops_e[k]:10d}\n")


if __name__ == "__main__":
    main()
