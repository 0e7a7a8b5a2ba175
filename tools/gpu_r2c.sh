#!/bin/bash
# round 2, call C: full -m gpu suite with the full-size exact-stream tests, ncu captures of the three instantiations
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2c_pytest.log
timeout 300 python tools/ab_probe.py --variants 4 --tag r2c > $O/r2c_ab.jsonl 2> $O/r2c_ab.err
for spec in "10 3840 2160 32 book1" "0 1920 1080 64 scene0" "8 1024 1024 32 scene8" "9 1920 1080 16 scene9"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o $O/r2c_hq_$5 -f \
     python bench.py --steps 1 --warmup 0 --scene $1 --width $2 --height $3 --spp $4 --no-cpu-baseline --no-e2e > $O/r2c_ncu_$5.log 2>&1
done
tail -3 $O/r2c_pytest.log; cat $O/r2c_ab.jsonl
