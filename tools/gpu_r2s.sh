#!/bin/bash
# round 2, call S: axis-aligned box fast path; scene 9 block size / leaf period sweep on the new kernel
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "exact_stream or tiles or slab_tested or bvh_modes" > $O/r2s_pytest.log 2>&1; tail -5 $O/r2s_pytest.log
CASES="7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag aabox > $O/r2s_ab.jsonl 2> $O/r2s_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases "9:1920x1080x32" --threads 576,512 --flags 0,0x10,0x20 --tag sweep9 >> $O/r2s_ab.jsonl 2>> $O/r2s_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases "9:1920x1080x32" --flags 0x10,0x20,0x400 --tag sweep9 >> $O/r2s_ab.jsonl 2>> $O/r2s_ab.err
cat $O/r2s_ab.jsonl | cut -c1-250
tail -3 $O/r2s_ab.err
