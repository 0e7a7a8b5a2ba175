#!/bin/bash
# round 2, call X: final-state validation: full -m gpu suite, smoke(), bench (+ reference arm), launch list, ncu summaries
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2x_pytest.log
grep -E "^FAILED|^E  " $O/r2x_pytest.log | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2x_smoke.log 2>&1; tail -2 $O/r2x_smoke.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2x_bench_reference.json 2> $O/r2x_bench_reference.err
timeout 900 python bench.py --steps 3 --warmup 3 > $O/r2x_bench.json 2> $O/r2x_bench.err; tail -5 $O/r2x_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2x_launches.csv \
    python bench.py --steps 2 --warmup 3 --spp 32 --no-cpu-baseline --no-configs > $O/r2x_launches.log 2>&1
LIB=$PWD/raytracinginoneweekendincuda_b200/librt_b200.so
for spec in "10 3840 2160 32 book1 RenderHitQueueILi0ELb1ELb0E RenderHitQueue<0,1,0>" "0 1920 1080 64 scene0 RenderHitQueueILi9ELb1ELb0E RenderHitQueue<9,1,0>" "8 1024 1024 32 scene8 RenderHitQueueILi6ELb1ELb0E RenderHitQueue<6,1,0>" "9 1920 1080 16 scene9 RenderHitQueueILi31ELb0ELb0E RenderHitQueue<31,0,0>"; do
  set -- $spec
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:RenderHitQueue -c 1 -o /tmp/r2x_$5 -f \
     python bench.py --steps 1 --warmup 0 --scene $1 --width $2 --height $3 --spp $4 --no-cpu-baseline --no-e2e --no-configs > $O/r2x_ncu_$5.log 2>&1
  python tools/ncu_summary.py /tmp/r2x_$5.ncu-rep $LIB $6 $O/r2x_hq_$5 "$7"
done
timeout 600 ncu --set full --clock-control none -k regex:ReduceResolve -c 1 -o /tmp/r2x_resolve -f python bench.py --steps 1 --warmup 0 --spp 8 --no-cpu-baseline --no-configs > $O/r2x_ncu_resolve.log 2>&1
ncu -i /tmp/r2x_resolve.ncu-rep --page details --csv > $O/r2x_resolve_details.csv 2>/dev/null
tail -3 $O/r2x_pytest.log; head -c 2500 $O/r2x_bench.json
