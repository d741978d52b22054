#!/usr/bin/env python
"""Table of the A/B timings written by `scripts/ab_variants.sh time` (gpurun_out/ab_<library>_<pass>.log): one row per
operation of scripts/gpu_probe.py at 2048^2, one column per library, best of the passes in microseconds and the ratio to
the product library (ratio < 1: the variant is faster).  `python scripts/ab_report.py [gpurun_out]`."""
import glob
import os
import re
import sys

d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
times = {}                                              # library -> operation -> best time
for path in sorted(glob.glob(os.path.join(d, "ab_*_[0-9].log"))):
    lib = re.sub(r"^ab_|_[0-9]\.log$", "", os.path.basename(path))
    size = None
    for line in open(path):
        m = re.match(r"--- N=(\d+)", line)
        if m:
            size = int(m.group(1))
            continue
        m = re.match(r"\s*(.+?)\s+([0-9.]+) us\s*$", line)
        if m and size == 2048:
            op, t = m.group(1).strip(), float(m.group(2))
            cur = times.setdefault(lib, {})
            cur[op] = min(t, cur.get(op, t))
if not times:
    sys.exit(f"no ab_*.log files under {d}")
base = "libxlprop"
libs = [base] + sorted(k for k in times if k != base)
ops = list(times[libs[0]].keys())
print(" " * 18 + "".join(f"{l.replace('libxlprop_', '')[:22]:>24s}" for l in libs))
for op in ops:
    row = f"{op:18s}"
    for l in libs:
        t = times.get(l, {}).get(op)
        ref = times.get(base, {}).get(op)
        row += f"{'':>24s}" if t is None else f"{t:14.1f} ({t / ref:5.3f})" + " "
    print(row)
for l in libs[1:]:
    p = os.path.join(d, f"ab_{l}_parity.log")
    if os.path.exists(p):
        print(f"parity {l}: {open(p).read().strip().splitlines()[-1]}")
