#!/bin/bash
# One gpurun call: GPU tests, the two bench arms, divergence probe, ncu launch list + full capture.
# usage: gpurun --timeout 1800 -- bash tools/gpu_round.sh [tag]
TAG=${1:-r1}
O=gpurun_out
mkdir -p $O
cp raytracinginoneweekendincuda_b200/librt_b200.so $O/librt_$TAG.so
tar czf $O/src_$TAG.tgz raytracinginoneweekendincuda_b200/csrc include
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu_$TAG.log
python bench.py > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"; cat $O/bench_$TAG.json; tail -3 $O/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_$TAG.json 2>> $O/bench_$TAG.err; cat $O/bench_ref_$TAG.json
python tools/gpu_probe.py diverge > $O/diverge_$TAG.log 2>&1; tail -4 $O/diverge_$TAG.log
CMD="python bench.py --steps 2 --warmup 3 --spp 32 --no-cpu-baseline --no-e2e"
$CMD > $O/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_$TAG.csv $CMD > $O/ncu_launches_$TAG.log 2>&1
$CMD > $O/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k "regex:Render(Mega|Wave|HeadTail)" -s 3 -c 1 -f -o $O/prof_${TAG}_book1 $CMD > $O/ncu_full_$TAG.log 2>&1
tail -3 $O/ncu_full_$TAG.log
