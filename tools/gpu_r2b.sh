#!/bin/bash
# round 2, call B: parity after the traversal-loop work, A/B of the FFMA slab build, leaf-period and block-size sweeps
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2b_pytest.log
timeout 300 python tools/ab_probe.py --variants 3,4 --tag base > $O/r2b_ab.jsonl 2> $O/r2b_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_ffma.so timeout 300 python tools/ab_probe.py --variants 3,4 --tag ffma >> $O/r2b_ab.jsonl 2>> $O/r2b_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases 10:3840x2160x64 --flags 0x0,0x10,0x20 --threads 0,736,704,640 --tag sweep >> $O/r2b_ab.jsonl 2>> $O/r2b_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases 10:3840x2160x64,9:1920x1080x32 --max-leaf 1,2,3,4 --tag leaf >> $O/r2b_ab.jsonl 2>> $O/r2b_ab.err
tail -3 $O/r2b_pytest.log; cat $O/r2b_ab.jsonl
