#!/bin/bash
# round 2, call P: hoisted items through the walk loop's leaf step (RT_HQ_STACKED_HOIST 1 = product, 0, 2)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "exact_stream or bvh_modes or slab_tested or hoisting or tiles or hit_queue" > $O/r2p_pytest.log 2>&1; tail -5 $O/r2p_pytest.log
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag hoist1 > $O/r2p_ab.jsonl 2> $O/r2p_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_hoist0.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag hoist0 >> $O/r2p_ab.jsonl 2>> $O/r2p_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_hoist2.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag hoist2 >> $O/r2p_ab.jsonl 2>> $O/r2p_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases "9:1920x1080x32" --max-leaf 1,2,4 --tag hoist1_leaf >> $O/r2p_ab.jsonl 2>> $O/r2p_ab.err
cat $O/r2p_ab.jsonl | cut -c1-250
tail -3 $O/r2p_ab.err
