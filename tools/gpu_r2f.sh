#!/bin/bash
# round 2, call F (2 GPUs): the full -m gpu suite incl. the two-device tests; single-process C-ABI bench and the torchrun
# bench with N-rank parity; Cornell-class instantiation timings
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi -L > $O/r2f_smi.txt
timeout 1800 python -m pytest tests -m gpu -q > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r2f_pytest.log
grep -E "^FAILED|^E  " $O/r2f_pytest.log | head -30
timeout 300 python tools/ab_probe.py --variants 4 --cases 7:1024x1024x64,8:1024x1024x64,9:1920x1080x32,10:3840x2160x64 --tag box > $O/r2f_ab.jsonl 2> $O/r2f_ab.err
timeout 600 python bench.py --gpus 2 --single-process --steps 3 --warmup 3 --no-configs > $O/r2f_bench_sp2.json 2> $O/r2f_bench_sp2.err; tail -3 $O/r2f_bench_sp2.err
timeout 600 python bench.py --gpus 2 --single-process --upload-flags 2 --steps 3 --warmup 3 --no-configs --no-e2e > $O/r2f_bench_sp2_nccl.json 2> $O/r2f_bench_sp2_nccl.err; tail -3 $O/r2f_bench_sp2_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r2f_bench_tr2.json 2> $O/r2f_bench_tr2.err; tail -3 $O/r2f_bench_tr2.err
timeout 300 ./raytracinginoneweekendincuda_b200/rt_cli --scene 10 --width 3840 --height 2160 --spp 256 --gpus 2 --p6 --out /tmp/a.ppm 2> $O/r2f_cli.txt; cat $O/r2f_cli.txt
tail -3 $O/r2f_pytest.log; cat $O/r2f_ab.jsonl | cut -c1-250
for f in sp2 sp2_nccl tr2; do python - <<PY
import json
try:
    d=json.load(open("$O/r2f_bench_$f.json"))
    print("$f", round(d["value"]), d["ms_per_step"], d.get("nrank_parity"), d.get("per_rank_kernel_ms"), (d.get("e2e") or {}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
