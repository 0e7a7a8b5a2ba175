#!/usr/bin/env python3
"""Attribute an ncu source-page CSV (SASS rows) to CUDA source lines.

usage: ncu_lines.py <report.ncu-rep> <library.so> <kernel-substring> [top] [srcdir]
Joins `ncu --page source --csv` (per-SASS-instruction samples / executed
instructions / active threads) with `nvdisasm -gi` line info by instruction
order.  Every SASS instruction carries its inline chain (innermost first);
three tables are printed: by outermost line (the call site in the kernel), by
innermost line, and by the function-level call site one level below the kernel.
`srcdir` = directory holding the sources the library was built from (default:
the working tree)."""
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
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
srcdir = sys.argv[5] if len(sys.argv) > 5 else ROOT
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = max((os.path.join(tmp, f) for f in os.listdir(tmp)), key=os.path.getsize)
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and kern in l and l.rstrip().endswith(":"))
chains = []  # per instruction: tuple of (file, line), innermost first
cur = []
fresh = True
for l in dis[start + 1:]:
    if l.startswith("//-----") or l.startswith("\t.section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            cur = []
            fresh = False
        cur.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        chains.append(tuple(cur) if cur else (("?", 0),))
        fresh = True
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(csvtxt)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ci = {n: i for i, n in enumerate(hdr)}
print(f"sass rows {len(data)}  disasm instrs {len(chains)}")
n = min(len(data), len(chains))


def f(x):
    try:
        return float(x)
    except ValueError:
        return 0.0


src_cache = {}


def src(key):
    fn, ln = key
    for base in ("raytracinginoneweekendincuda_b200/csrc", "include", "."):
        p = os.path.join(srcdir, base, fn)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            if 0 < ln <= len(src_cache[p]):
                return src_cache[p][ln - 1].strip()
    return ""


def table(title, keyfn):
    agg = defaultdict(lambda: [0.0, 0.0, 0.0])
    for r, ch in zip(data[:n], chains[:n]):
        a = agg[keyfn(ch)]
        a[0] += f(r[ci["# Samples"]])
        a[1] += f(r[ci["Instructions Executed"]])
        a[2] += f(r[ci["Thread Instructions Executed"]])
    ts = sum(a[0] for a in agg.values())
    ti = sum(a[1] for a in agg.values())
    tt = sum(a[2] for a in agg.values())
    print(f"\n== {title}: samples {ts:.0f}  warp-instr {ti:.3e}  thread-instr {tt:.3e}  avg active {tt / max(ti, 1):.2f}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{a[0] / ts * 100:5.1f}% samples {a[1] / ti * 100:5.1f}% instr  act {a[2] / max(a[1], 1):5.1f}  "
              f"{key[0]}:{key[1]:<4} {src(key)[:100]}")


table("by kernel-level line (outermost)", lambda ch: ch[-1])
table("one level below the kernel", lambda ch: ch[-2] if len(ch) > 1 else ch[-1])
table("by innermost line", lambda ch: ch[0])
