#!/bin/bash
# round 2, call L: eager lens-disk sampling; 2 vs 4 eagerly pre-tested ball candidates
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,10:3840x2160x256,0:1920x1080x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag eager2_disk > $O/r2l_ab.jsonl 2> $O/r2l_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_eager4.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag eager4_disk >> $O/r2l_ab.jsonl 2>> $O/r2l_ab.err
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "exact_stream or hit_queue or deterministic or megakernel" > $O/r2l_pytest.log 2>&1; tail -3 $O/r2l_pytest.log
cat $O/r2l_ab.jsonl | cut -c1-250
