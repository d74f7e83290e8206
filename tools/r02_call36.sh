#!/bin/bash
# Round 2, GPU call 36 (4 GPUs): the bench at N = 4 on the final code (NB_HOST_THREADS = cores / 4 per rank).
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $O/c36_bench_n4.json 2> $O/c36_bench_n4.err
tail -c 200 $O/c36_bench_n4.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c36_bench_n4.json").read().strip().splitlines()[-1])
    print("N=4", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")}, d["e2e"]["ms_per_step"])
except Exception as e: print("n4 failed", e)
PY
