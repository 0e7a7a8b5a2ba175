#!/bin/bash
# round 2, call O: slab-tested boxes (DevBox) vs quad-by-quad (RT_UPLOAD_NO_BOXES = 8); spp dependence of scene 9
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "exact_stream or bvh_modes or slab_tested or hoisting or tiles or full_size" > $O/r2o_pytest.log 2>&1; tail -5 $O/r2o_pytest.log
CASES="7:1024x1024x64,8:1024x1024x64,9:1920x1080x32,9:3840x2160x8,9:960x540x128,9:480x270x512"
timeout 600 python tools/ab_probe.py --variants 4 --cases $CASES --upload-flags 0,8 --tag boxes > $O/r2o_ab.jsonl 2> $O/r2o_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases "10:3840x2160x64,0:1920x1080x64" --tag boxes >> $O/r2o_ab.jsonl 2>> $O/r2o_ab.err
cat $O/r2o_ab.jsonl | cut -c1-250
tail -3 $O/r2o_ab.err
