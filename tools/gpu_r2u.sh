#!/bin/bash
# round 2, call U: ncu of the Cornell-class kernel (scenes 7, 8) and Book 1 after the single-exit rejection loop
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
LIB=$PWD/raytracinginoneweekendincuda_b200/librt_b200.so
for spec in "8 1024 1024 32 scene8 RenderHitQueueILi6ELb1ELb0E" "10 3840 2160 32 book1 RenderHitQueueILi0ELb1ELb0E"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o /tmp/r2u_$5 -f \
     python bench.py --steps 1 --warmup 0 --scene $1 --width $2 --height $3 --spp $4 --no-cpu-baseline --no-e2e --no-configs > $O/r2u_ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/r2u_$5.ncu-rep $LIB $6 $O/r2u_hq_$5
  python tools/ncu_sass.py /tmp/r2u_$5.ncu-rep 0.05 > $O/r2u_hq_$5_sass_full.txt
done
ls -la $O | grep r2u
