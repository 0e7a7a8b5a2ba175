#!/bin/bash
# usage: gpurun -- bash tools/gpu_prof.sh <tag> <bench args...>   (ncu --set full of RenderMega on a short bench run)
TAG=$1; shift
O=gpurun_out
mkdir -p $O
cp raytracinginoneweekendincuda_b200/librt_b200.so $O/librt_$TAG.so
tar czf $O/src_$TAG.tgz raytracinginoneweekendincuda_b200/csrc include
CMD="python bench.py --steps 2 --warmup 3 --spp 32 --no-cpu-baseline --no-e2e $*"
$CMD > $O/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:Render(Mega|Wave|HeadTail)" -s 3 -c 1 -f -o $O/prof_$TAG $CMD > $O/ncu_full_$TAG.log 2>&1
tail -2 $O/plain_$TAG.log | cut -c1-300; tail -3 $O/ncu_full_$TAG.log
