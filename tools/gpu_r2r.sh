#!/bin/bash
# round 2, call R: single-exit rejection loop (RT_BALL_STRUCTURED 1 = product vs 0); walk loop unrolled for every kernel
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "exact_stream or tiles or hit_queue or deterministic" > $O/r2r_pytest.log 2>&1; tail -5 $O/r2r_pytest.log
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag ball1 > $O/r2r_ab.jsonl 2> $O/r2r_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_ball0.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag ball0 >> $O/r2r_ab.jsonl 2>> $O/r2r_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_unroll2.so timeout 300 python tools/ab_probe.py --variants 4 --cases "9:1920x1080x32" --tag unroll2 >> $O/r2r_ab.jsonl 2>> $O/r2r_ab.err
cat $O/r2r_ab.jsonl | cut -c1-250
tail -3 $O/r2r_ab.err
