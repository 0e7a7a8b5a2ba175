#!/bin/bash
# BASELINE configs[4] (Book 2 final, 3840x2160, 10000 spp) on N GPUs: usage gpu_config5_ngpu.sh N
N=${1:-8}
set -x
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --scene 9 --width 3840 --height 2160 --spp 10000 --steps 1 --warmup 1 --no-configs --no-cpu-baseline --no-e2e $EXTRA > $O/r2_config5_n$N.json 2> $O/r2_config5_n$N.err
tail -2 $O/r2_config5_n$N.err
python - <<PY
import json
d=json.load(open("$O/r2_config5_n$N.json"))
print($N, round(d["value"]), round(d["ms_per_step"],1), d.get("nrank_parity"), d.get("per_rank_kernel_ms"))
PY
