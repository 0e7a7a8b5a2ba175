#!/usr/bin/env python3
"""Boil an `ncu --set full` report down to the small text files profiles/ keeps (run ON the GPU box: reports are
tens of MB and gpurun brings back at most 64 MiB).

usage: ncu_summary.py <report.ncu-rep> <library.so> <mangled-kernel-substring> <out-prefix> [traffic-key]
writes <out-prefix>_raw_selected.txt   pipe utilisations, stall reasons, DRAM bytes, occupancy, registers
       <out-prefix>_source_lines.txt   tools/ncu_lines.py: % of issued warp instructions + active lanes per CUDA line
       <out-prefix>_hot_sass.txt       tools/ncu_sass.py: the hot SASS
       <out-prefix>_details.csv        ncu details page
and, with traffic-key, merges {"<key>": {"dram_bytes": read+write, ...}} into gpurun_out/r2_traffic.json."""
import csv
import io
import json
import os
import re
import subprocess
import sys

rep, lib, kern, prefix = sys.argv[1:5]
key = sys.argv[5] if len(sys.argv) > 5 else None
HERE = os.path.dirname(os.path.abspath(__file__))
KEEP = re.compile(r"^(dram__bytes_(read|write)\.sum$|gpu__time_duration\.sum|gpu__dram_throughput|l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|"
                  r"launch__(registers_per_thread|block_size|grid_size|shared_mem_per_block_dynamic|occupancy_limit)|"
                  r"sm__inst_executed_pipe_[a-z0-9_]+\.avg\.pct_of_peak_sustained_active$|sm__warps_active\.avg\.pct|"
                  r"smsp__average_warps_issue_stalled_[a-z_]+_per_issue_active\.ratio$|smsp__inst_executed\.sum$|"
                  r"smsp__issue_active\.avg\.pct|smsp__thread_inst_executed_per_inst_executed\.ratio$|sm__cycles_elapsed\.max$|"
                  r"l1tex__t_sector_hit_rate\.pct$|lts__t_sector_hit_rate\.pct$|l1tex__throughput\.avg\.pct|lts__throughput\.avg\.pct|"
                  r"sm__throughput\.avg\.pct|smsp__inst_executed_op_shared|local_(load|store))")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
sel = {}
with open(prefix + "_raw_selected.txt", "w") as f:
    f.write(f"# {os.path.basename(rep)}: kernel {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else kern}\n")
    for n, u, v in sorted(zip(hdr, units, vals)):
        if KEEP.search(n):
            f.write(f"{n} {u} {v}\n")
            sel[n] = v
for tool, out, extra in (("ncu_lines.py", "_source_lines.txt", [lib, kern, "70"]), ("ncu_sass.py", "_hot_sass.txt", ["0.25"])):
    txt = subprocess.run([sys.executable, os.path.join(HERE, tool), rep] + extra, capture_output=True, text=True)
    open(prefix + out, "w").write(txt.stdout + (("\n[stderr]\n" + txt.stderr[-2000:]) if txt.returncode else ""))
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
open(prefix + "_details.csv", "w").write(det)
if key:
    def num(x):
        return float(x.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for n in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        u = units[hdr.index(n)]
        tot += num(vals[hdr.index(n)]) * mult.get(u, 1)
    path = os.path.join(os.path.dirname(prefix), "r2_traffic.json")
    cur = json.load(open(path)) if os.path.exists(path) else {}
    cur[key] = {"dram_bytes": int(tot), "report": os.path.basename(rep),
                "duration_ms": sel.get("gpu__time_duration.sum"), "note": "one launch, ncu --set full --clock-control none"}
    json.dump(cur, open(path, "w"), indent=1)
print("summarised", rep, "->", prefix)
