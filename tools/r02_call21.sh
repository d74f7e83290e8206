#!/bin/bash
# Round 2, GPU call 21 (4 GPUs): bench at N = 4, index sharded + all-gathered vs replicated per rank.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
for mode in sharded replicated; do
NECAT_B200_INDEX=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > $O/c21_bench_n4_$mode.json 2> $O/c21_bench_n4_$mode.err
tail -c 200 $O/c21_bench_n4_$mode.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c21_bench_n4_$mode.json").read().strip().splitlines()[-1])
    print("N=4 $mode", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")}, d["e2e"]["ms_per_step"])
except Exception as e: print("n4 failed", e)
PY
done
