#!/bin/bash
# multi-GPU validation (run with gpurun --gpus N, N = 2 or 8): two-device tests, single-process C-ABI bench, torchrun bench
# usage: bash tools/gpu_final_ngpu.sh N TAG
N=${1:-2}; TAG=${2:-r2n}
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi -L > $O/${TAG}_smi.txt
if [ "$N" = "2" ]; then
  timeout 1800 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
  grep -E "^FAILED|^E  " $O/${TAG}_pytest.log | head -30; tail -2 $O/${TAG}_pytest.log
else
  timeout 900 python -m pytest tests/test_multi_device_gpu.py -m gpu -q > $O/${TAG}_pytest.log 2>&1; tail -2 $O/${TAG}_pytest.log
fi
timeout 600 python bench.py --gpus $N --single-process --steps 3 --warmup 3 --no-configs > $O/${TAG}_bench_sp.json 2> $O/${TAG}_bench_sp.err; tail -3 $O/${TAG}_bench_sp.err
timeout 600 python bench.py --gpus $N --single-process --upload-flags 2 --steps 3 --warmup 3 --no-configs --no-e2e > $O/${TAG}_bench_sp_nccl.json 2> $O/${TAG}_bench_sp_nccl.err; tail -3 $O/${TAG}_bench_sp_nccl.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 3 --warmup 3 > $O/${TAG}_bench_tr.json 2> $O/${TAG}_bench_tr.err; tail -3 $O/${TAG}_bench_tr.err
timeout 300 ./raytracinginoneweekendincuda_b200/rt_cli --scene 10 --width 3840 --height 2160 --spp 1024 --gpus $N --p6 --out /tmp/a.ppm 2> $O/${TAG}_cli.txt; cat $O/${TAG}_cli.txt
for f in sp sp_nccl tr; do python - <<PY
import json
try:
    d=json.load(open("$O/${TAG}_bench_$f.json"))
    print("$f", round(d["value"]), round(d["ms_per_step"],2), d.get("nrank_parity"), d.get("nrank_parity_detail",{}).get("max_rel_diff"), d.get("nrank_parity_detail",{}).get("bit_identical_fraction"), d.get("per_rank_kernel_ms"), (d.get("e2e") or {}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
