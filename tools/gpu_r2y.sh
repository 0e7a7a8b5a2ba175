#!/bin/bash
# round 2, call Y: 3 / 4 box steps per leaf step in the unrolled walk loop (small kernels)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag steps2 > $O/r2y_ab.jsonl 2> $O/r2y_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_steps3.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag steps3 >> $O/r2y_ab.jsonl 2>> $O/r2y_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_steps4.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag steps4 >> $O/r2y_ab.jsonl 2>> $O/r2y_ab.err
cat $O/r2y_ab.jsonl | cut -c1-250
