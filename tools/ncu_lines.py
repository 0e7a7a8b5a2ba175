#!/usr/bin/env python3
"""Attribute an ncu source-page CSV (SASS rows) to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <library.so> <kernel-substring> [top]
Joins `ncu --page source --csv` (per-SASS-instruction samples / executed
instructions / active threads) with `nvdisasm -gi` line info by instruction
order, then prints the hottest source lines and per-function totals."""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict

rep, lib, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel section
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
lines = []
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
print(f"sass rows {len(data)}  disasm instrs {len(lines)}")
n = min(len(data), len(lines))


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


agg = defaultdict(lambda: [0.0, 0.0, 0.0])
src_cache = {}
for r, key in zip(data[:n], lines[:n]):
    a = agg[key]
    a[0] += f(r[ci["# Samples"]])
    a[1] += f(r[ci["Instructions Executed"]])
    a[2] += f(r[ci["Thread Instructions Executed"]])
ts = sum(a[0] for a in agg.values())
ti = sum(a[1] for a in agg.values())
tt = sum(a[2] for a in agg.values())
print(f"samples {ts:.0f}  warp-instr {ti:.3e}  thread-instr {tt:.3e}  avg active {tt / ti:.2f}")


def src(key):
    fn, ln = key
    for base in ("raytracinginoneweekendincuda_b200/csrc", "include"):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), base, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < ln <= len(src_cache[p]):
                return src_cache[p][ln - 1].strip()
    return ""


for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{a[0] / ts * 100:5.1f}% stall-samples {a[1] / ti * 100:5.1f}% instr  act {a[2] / max(a[1], 1):5.1f}  {key[0]}:{key[1]:<4} {src(key)[:90]}")
