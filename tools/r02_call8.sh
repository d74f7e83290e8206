#!/bin/bash
# Round 2, GPU call 8 (2 GPUs): the volume-sharded pipeline test, bench at N=2 and N=1 (with parity + e2e_exe).
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/c8_pytest_multi.log 2>&1
tail -5 $O/c8_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > $O/c8_bench_n2.json 2> $O/c8_bench_n2.err
tail -c 300 $O/c8_bench_n2.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c8_bench_n2.json").read().strip().splitlines()[-1])
    print("N=2", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")})
except Exception as e: print("n2 failed", e)
PY
timeout 900 python bench.py > $O/c8_bench_n1.json 2> $O/c8_bench_n1.err
tail -c 300 $O/c8_bench_n1.err
python - <<PY
import json
d = json.loads(open("$O/c8_bench_n1.json").read().strip().splitlines()[-1])
print("N=1", {k: d.get(k) for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates", "parity", "e2e_exe", "cpu_baseline", "roofline_seed_vote")})
print(d["roofline"])
PY
