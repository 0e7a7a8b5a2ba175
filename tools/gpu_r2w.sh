#!/bin/bash
# round 2, call W: BVH4 (RT_BVH4=1 build: collapsed four-wide nodes) against the shipping binary tree
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,0:1920x1080x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag bvh2 > $O/r2w_ab.jsonl 2> $O/r2w_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases "10:3840x2160x64" --threads 704 --tag bvh2_704 >> $O/r2w_ab.jsonl 2>> $O/r2w_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_bvh4.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag bvh4 >> $O/r2w_ab.jsonl 2>> $O/r2w_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_bvh4.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "exact_stream_parity_every_scene or config1" > $O/r2w_pytest.log 2>&1; tail -3 $O/r2w_pytest.log
cat $O/r2w_ab.jsonl | cut -c1-260
# instruction counts / lanes of the two builds on Book 1 (ncu, one launch each)
for v in bvh2 bvh4; do
  L=""; [ $v = bvh4 ] && L=$PWD/tools/variants/librt_bvh4.so
  RT_B200_LIBRARY=$L timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:RenderHitQueue -c 1 --csv --log-file $O/r2w_ncu_$v.csv \
     python bench.py --steps 1 --warmup 0 --scene 10 --width 3840 --height 2160 --spp 32 --no-cpu-baseline --no-e2e --no-configs > $O/r2w_ncu_$v.log 2>&1
  tail -6 $O/r2w_ncu_$v.csv | cut -d, -f5,13-
done
