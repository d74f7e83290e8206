#!/bin/bash
# Round 2, GPU call 30 (2 GPUs): the sharded pipeline tests and the bench at N = 2 on the final code (index / mapping kernels
# changed since call 21; NB_HOST_THREADS per rank).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > $O/c30_pytest_multi.log 2>&1
tail -3 $O/c30_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/c30_bench_n2.json 2> $O/c30_bench_n2.err
tail -c 300 $O/c30_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c30_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")}, d["e2e"]["ms_per_step"])
except Exception as e: print("n2 failed", e)
PY
