#!/bin/bash
# Round 2, GPU call 25 (8 GPUs): the bench at N = 8.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > $O/c25_bench_n8.json 2> $O/c25_bench_n8.err
tail -c 200 $O/c25_bench_n8.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c25_bench_n8.json").read().strip().splitlines()[-1])
    print("N=8", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")}, d["e2e"]["ms_per_step"])
except Exception as e: print("n8 failed", e)
PY
