#!/bin/bash
# round 2, call V: quad interior test as two triple products, compare-and-select for the FP64 min/max of the box media
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_importance.py -m gpu -q -x -k "exact_stream or tiles or slab_tested or bvh_modes or importance or hit_queue" > $O/r2v_pytest.log 2>&1; tail -5 $O/r2v_pytest.log
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag triple > $O/r2v_ab.jsonl 2> $O/r2v_ab.err
cat $O/r2v_ab.jsonl | cut -c1-250
