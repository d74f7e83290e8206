#!/bin/bash
# Round 2, GPU call 14: consensus in slices (serialised link kernels, chains beside them), dynamic replay scheduling, map trace.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cns.py -x -q > $O/c14_pytest.log 2>&1
tail -2 $O/c14_pytest.log
for env in "A=1" "NB_CNS_SLICES=2" "NB_CNS_SLICES=4" "NB_CNS_SLICES=5 NB_CNS_LONG_PCT=6" "NB_CNS_SLICES=4 NB_CNS_LONG_PCT=12"; do
env $env NB_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > $O/c14_tmp.json 2> $O/c14_tmp.err
python - "$env" <<PY
import json, sys
try:
    d = json.loads(open("$O/c14_tmp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms/step %.1f e2e %.1f | ext %.1f cns %.1f | wall %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms"]["extend_ms"], d["stage_ms"]["consensus_ms"], {k: round(v) for k, v in d["wall_ms"].items()}))
except Exception as e: print(sys.argv[1], "FAILED", e, open("$O/c14_tmp.err").read()[-300:])
PY
grep -E "nb_consensus_job" $O/c14_tmp.err | tail -1 | cut -c1-600
grep -E "nb_map_range" $O/c14_tmp.err | tail -1 | cut -c1-600
done
cp $O/c14_tmp.json $O/c14_bench_last.json
