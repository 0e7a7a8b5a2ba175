#!/bin/bash
# round 2, call A: parity of the new kernel, A/B timings, bench, ncu capture of the hit-queue kernel
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2a_pytest.log
timeout 600 python tools/ab_probe.py --variants 3,4 --upload-flags 0,1 > $O/r2a_ab.jsonl 2> $O/r2a_ab.err
timeout 600 python bench.py --steps 3 --warmup 3 > $O/r2a_bench.json 2> $O/r2a_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o $O/r2a_hq_book1 -f \
   python bench.py --steps 1 --warmup 0 --spp 32 --no-cpu-baseline --no-e2e > $O/r2a_ncu.log 2>&1
tail -3 $O/r2a_pytest.log; cat $O/r2a_ab.jsonl; cat $O/r2a_bench.json | head -c 1500
