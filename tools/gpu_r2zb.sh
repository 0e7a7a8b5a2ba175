#!/bin/bash
# round 2, call ZB: SAH cost model (primitive cost against the node cost) and leaf size
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,0:1920x1080x64,9:1920x1080x32"
: > $O/r2zb_ab.jsonl
for sc in 0.5 1 2 4; do
  RT_SAH_ISECT_SCALE=$sc timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --max-leaf 1,2,4 --tag isect$sc >> $O/r2zb_ab.jsonl 2>> $O/r2zb_ab.err
done
cat $O/r2zb_ab.jsonl | cut -c1-250
