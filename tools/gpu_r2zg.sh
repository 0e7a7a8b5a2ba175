#!/bin/bash
# round 2, call ZG: ncu of the feature-complete kernel at the final state (Perlin table in shared memory)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
LIB=$PWD/raytracinginoneweekendincuda_b200/librt_b200.so
timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o /tmp/r2zg_scene9 -f \
   python bench.py --steps 1 --warmup 0 --scene 9 --width 1920 --height 1080 --spp 16 --no-cpu-baseline --no-e2e --no-configs > $O/r2zg_ncu_scene9.log 2>&1
python tools/ncu_summary.py /tmp/r2zg_scene9.ncu-rep $LIB RenderHitQueueILi31ELb0ELb0E $O/r2zg_hq_scene9 "RenderHitQueue<31,0,0>"
grep -E "time_duration|issue_active.avg|thread_inst_executed_per|inst_executed.sum|long_scoreboard|no_instruction|l1tex__t_sector_hit|shared_mem_per_block" $O/r2zg_hq_scene9_raw_selected.txt
