#!/bin/bash
# round 2, call Q: ncu of the Cornell-class and feature-complete kernels after the slab-tested boxes
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
LIB=$PWD/raytracinginoneweekendincuda_b200/librt_b200.so
for spec in "7 1024 1024 32 scene7 RenderHitQueueILi6ELb1ELb0E RenderHitQueue<6,1,0>/scene7" "8 1024 1024 32 scene8 RenderHitQueueILi6ELb1ELb0E RenderHitQueue<6,1,0>" "9 1920 1080 16 scene9 RenderHitQueueILi31ELb0ELb0E RenderHitQueue<31,0,0>"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o /tmp/r2q_$5 -f \
     python bench.py --steps 1 --warmup 0 --scene $1 --width $2 --height $3 --spp $4 --no-cpu-baseline --no-e2e --no-configs > $O/r2q_ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/r2q_$5.ncu-rep $LIB $6 $O/r2q_hq_$5 "$7"
done
ls -la $O | grep r2q
