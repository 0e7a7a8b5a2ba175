#!/bin/bash
# round 2, call ZA: more warps per CTA at fewer registers (800 / 832 threads at 72 registers; 672 / 704 at 80 for the large kernels)
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
CASES="10:3840x2160x64,0:1920x1080x64,8:1024x1024x64,9:1920x1080x32"
timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag t768_640 > $O/r2za_ab.jsonl 2> $O/r2za_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_t800.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag t800_672 >> $O/r2za_ab.jsonl 2>> $O/r2za_ab.err
RT_B200_LIBRARY=$PWD/tools/variants/librt_t832.so timeout 300 python tools/ab_probe.py --variants 4 --cases $CASES --tag t832_704 >> $O/r2za_ab.jsonl 2>> $O/r2za_ab.err
cat $O/r2za_ab.jsonl | cut -c1-250
