#!/bin/bash
# round 2, call ZC: four eager ball candidates in the single-exit form
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag eager2s > $O/r2zc_ab.jsonl 2> $O/r2zc_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_ball4.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag eager4s >> $O/r2zc_ab.jsonl 2>> $O/r2zc_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_ball4.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "exact_stream_parity_every_scene" > $O/r2zc_pytest.log 2>&1; tail -2 $O/r2zc_pytest.log
cat $O/r2zc_ab.jsonl | cut -c1-250
