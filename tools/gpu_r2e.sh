#!/bin/bash
# round 2, call E: new hoisting policy (media + tiny scenes), noinline HitMedium A/B, full gpu test suite
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2e_pytest.log
grep -E "^FAILED|^E  " $O/r2e_pytest.log | head -30
CASES="10:3840x2160x64,0:1920x1080x64,7:1024x1024x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --upload-flags 0,1 --tag noinline > $O/r2e_ab.jsonl 2> $O/r2e_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_medinl.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag inline >> $O/r2e_ab.jsonl 2>> $O/r2e_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases 8:1024x1024x64,9:1920x1080x32 --threads 512,576,640 --tag threads >> $O/r2e_ab.jsonl 2>> $O/r2e_ab.err
tail -3 $O/r2e_pytest.log; cat $O/r2e_ab.jsonl | cut -c1-260
