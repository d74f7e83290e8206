#!/bin/bash
# Round 2, GPU call 22: the whole -m gpu suite, __graft_entry__.smoke(), and the bench line (fewer collectives per level in k_cns_dp).
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/c22_pytest_gpu.log 2>&1
tail -4 $O/c22_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
NB_TRACE=1 timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/c22_bench.json 2> $O/c22_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c22_bench.json").read().strip().splitlines()[-1])
    print("bench", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, d["e2e"]["ms_per_step"])
except Exception as e: print("bench failed", e)
PY
grep -E "nb_consensus_job" $O/c22_bench.err | head -3 | tail -1 | cut -c1-600
