#!/bin/bash
# round 2, call I: whole MakeBox lists as single BVH items, node table staged for large scenes, sphere-boundary media
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2i_pytest.log
grep -E "^FAILED|^E  " $O/r2i_pytest.log | head -30
timeout 300 python tools/ab_probe.py --variants 4 --cases 9:1920x1080x32,9:3840x2160x8 --upload-flags 0,4 --flags 0x0,0x400 --threads 0,576,512 --tag s9 > $O/r2i_ab.jsonl 2> $O/r2i_ab.err
timeout 300 python tools/ab_probe.py --variants 4 --cases 7:1024x1024x64,8:1024x1024x64,10:3840x2160x64,0:1920x1080x64 --upload-flags 0,4 --tag others >> $O/r2i_ab.jsonl 2>> $O/r2i_ab.err
tail -3 $O/r2i_pytest.log; cat $O/r2i_ab.jsonl | cut -c1-250
