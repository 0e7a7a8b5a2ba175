#!/bin/bash
# round 2, call ZE: two-bin hit queue (sphere hits / quad + medium hits) for the scenes with quads
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_importance.py -m gpu -q -x -k "exact_stream or tiles or hit_queue or deterministic or glass_metal or importance or sample_ranges" > $O/r2ze_pytest.log 2>&1; tail -4 $O/r2ze_pytest.log
CASES="5:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32,9:3840x2160x16"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --reps 4 --tag bins > $O/r2ze_ab.jsonl 2> $O/r2ze_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_nobins.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --reps 4 --tag nobins >> $O/r2ze_ab.jsonl 2>> $O/r2ze_ab.err
cat $O/r2ze_ab.jsonl | cut -c1-250
