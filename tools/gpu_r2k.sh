#!/bin/bash
# round 2, call K: eager two-candidate ball sampling A/B; leaf-turn period on the Cornell-class / feature-complete kernels
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,10:3840x2160x256,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag eager > $O/r2k_ab.jsonl 2> $O/r2k_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_noeager.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag noeager >> $O/r2k_ab.jsonl 2>> $O/r2k_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases 7:1024x1024x64,8:1024x1024x64,9:1920x1080x32,0:1920x1080x64 --flags 0x0,0x10,0x20,0x30 --tag leafperiod >> $O/r2k_ab.jsonl 2>> $O/r2k_ab.err
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "exact_stream or hit_queue or deterministic" > $O/r2k_pytest.log 2>&1; tail -3 $O/r2k_pytest.log
cat $O/r2k_ab.jsonl | cut -c1-250
