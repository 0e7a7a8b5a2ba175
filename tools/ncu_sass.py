#!/usr/bin/env python3
"""Hot SASS of a kernel from an ncu report: per instruction, executed warp instructions (in % of the kernel), average
active threads and stall samples, with the CUDA source line it belongs to.

usage: ncu_sass.py <report.ncu-rep> [min_percent=0.3] [--range lo hi]   (lo/hi = SASS row indices)
Rows below min_percent are folded into '...' markers so the loops stand out."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 0.3
rng = None
if "--range" in sys.argv:
    k = sys.argv.index("--range")
    rng = (int(sys.argv[k + 1]), int(sys.argv[k + 2]))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
tot = sum(float(r[ci["Instructions Executed"]]) for r in data)
tots = sum(float(r[ci["# Samples"]]) for r in data)
print(f"{len(data)} SASS rows, {tot:.3e} warp instructions, {tots:.0f} samples")
skipped = 0
for k, r in enumerate(data):
    if rng and not (rng[0] <= k <= rng[1]):
        continue
    ie = float(r[ci["Instructions Executed"]])
    pct = ie / tot * 100
    if pct < minp and not rng:
        skipped += 1
        continue
    if skipped:
        print(f"      ... {skipped} rows")
        skipped = 0
    act = float(r[ci["Thread Instructions Executed"]]) / max(ie, 1)
    sm = float(r[ci["# Samples"]]) / tots * 100
    print(f"{k:5d} {pct:5.2f}% act {act:5.1f} smp {sm:5.2f}%  {r[ci['Source']].strip()}")
