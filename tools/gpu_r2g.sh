#!/bin/bash
# round 2, call G: logf medium draws + Cornell-class kernel in-tree; walk-loop unroll A/B; bench
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2g_pytest.log
grep -E "^FAILED|^E  " $O/r2g_pytest.log | head -30
CASES="10:3840x2160x64,10:3840x2160x256,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag base > $O/r2g_ab.jsonl 2> $O/r2g_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_unroll.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag unroll >> $O/r2g_ab.jsonl 2>> $O/r2g_ab.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2g_bench.json 2> $O/r2g_bench.err; tail -5 $O/r2g_bench.err
tail -3 $O/r2g_pytest.log; cat $O/r2g_ab.jsonl | cut -c1-250
