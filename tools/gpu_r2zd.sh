#!/bin/bash
# round 2, call ZD: leaf period of the feature-complete kernel (flags 0x20 = every 4th step, 0x30 = every 8th); leaf prefetch
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="9:1920x1080x32,9:3840x2160x16,2:1920x1080x32,3:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --flags 0,0x20,0x30 --reps 4 --tag leafperiod > $O/r2zd_ab.jsonl 2> $O/r2zd_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_prefetch.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --flags 0,0x20 --reps 4 --tag prefetch >> $O/r2zd_ab.jsonl 2>> $O/r2zd_ab.err
cat $O/r2zd_ab.jsonl | cut -c1-250
