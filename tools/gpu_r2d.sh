#!/bin/bash
# round 2, call D: full -m gpu suite (new multi-device / progressive / CLI tests), bench with configs array,
# ncu captures of the three instantiations summarised on the box
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2d_pytest.log
grep -E "FAILED|Error|assert" $O/r2d_pytest.log | head -20
timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2d_bench.json 2> $O/r2d_bench.err; tail -5 $O/r2d_bench.err
LIB=$PWD/raytracinginoneweekendincuda_b200/librt_b200.so
for spec in "10 3840 2160 32 book1 RenderHitQueueILi0ELb1ELb0E RenderHitQueue<0,1,0>" "0 1920 1080 64 scene0 RenderHitQueueILi9ELb1ELb0E RenderHitQueue<9,1,0>" "8 1024 1024 32 scene8 RenderHitQueueILi31ELb1ELb0E RenderHitQueue<31,1,0>" "9 1920 1080 16 scene9 RenderHitQueueILi31ELb0ELb0E RenderHitQueue<31,0,0>"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o /tmp/r2d_$5 -f \
     python bench.py --steps 1 --warmup 0 --scene $1 --width $2 --height $3 --spp $4 --no-cpu-baseline --no-e2e --no-configs > $O/r2d_ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/r2d_$5.ncu-rep $LIB $6 $O/r2d_hq_$5 "$7" 
done
tail -3 $O/r2d_pytest.log; head -c 3000 $O/r2d_bench.json
