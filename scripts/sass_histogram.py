#!/usr/bin/env python
"""Per-kernel SASS opcode evidence for the Blackwell claims (tcgen05 / TMEM / TMA), from the in-tree shared library.

    python scripts/sass_histogram.py [out.txt]        (default: profiles/sass_r02.txt; needs cuobjdump, no GPU)

Mnemonics (sm_100a): UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), UTCBAR = tcgen05.commit -> mbarrier, LDTM / STTM =
tcgen05.ld / tcgen05.st (tensor memory <-> registers), UBLKCP = cp.async.bulk (1-D TMA copy), SYNCS = mbarrier ops,
FADD2 / FFMA2 = packed fp32, HMMA = mma.sync (register-operand tensor-core instruction: only the backward contractions of
the fp32 fused kernel use it, HMMA.1688.F32.TF32 in three passes per product -- csrc/fused.cu explains why tcgen05 cannot
serve that kernel), IMMA: none."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "globecom2020-resourceallocationgnn_b200", "libv2vgnn_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "FADD2", "FFMA2",
         "FFMA", "HMMA", "IMMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ACQBULK", "ELECT"]


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "sass_r02.txt")
    txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    rows = []
    for (name, cnt), dn in zip(kernels.items(), demangled):
        short = re.sub(r"\(.*", "", dn.replace("(anonymous namespace)::", "")).replace("void ", "").replace("v2v::", "")
        rows.append((short, sum(cnt.values()), cnt))
    cols = [w for w in WATCH if any(r[2][w] for r in rows)]
    with open(out_path, "w") as o:
        o.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: static instruction counts per kernel (scripts/sass_histogram.py)\n")
        o.write("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA 1-D), SYNCS = mbarrier\n")
        o.write(f"{'kernel':58s} {'total':>7s} " + " ".join(f"{c:>7s}" for c in cols) + "\n")
        for short, total, cnt in sorted(rows, key=lambda r: -r[1]):
            o.write(f"{short[:58]:58s} {total:7d} " + " ".join(f"{cnt[c]:7d}" for c in cols) + "\n")
        tot = collections.Counter()
        for _, _, cnt in rows:
            tot.update(cnt)
        o.write(f"{'ALL KERNELS':58s} {sum(tot.values()):7d} " + " ".join(f"{tot[c]:7d}" for c in cols) + "\n")
        legacy = tot["HMMA"] + tot["IMMA"]
        o.write(f"# legacy mma.sync instructions (HMMA/IMMA): {legacy}\n")
    print(open(out_path).read()[:3000])


if __name__ == "__main__":
    main()
