#!/bin/bash
# Round 2, GPU call 15: pipelined template groups (own streams and scratch, launches queued in group order), leaner map grouping.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cns.py tests/test_gpu_pmov.py -x -q > $O/c15_pytest.log 2>&1
tail -2 $O/c15_pytest.log
for env in "NB_CNS_GROUPS=1" "NB_CNS_GROUPS=2" "NB_CNS_GROUPS=3" "NB_CNS_GROUPS=4" "NB_CNS_GROUPS=2 NB_CNS_SLICES=2"; do
env $env NB_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > $O/c15_tmp.json 2> $O/c15_tmp.err
python - "$env" <<PY
import json, sys
try:
    d = json.loads(open("$O/c15_tmp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms/step %.1f e2e %.1f | ext %.1f cns %.1f (timed passes: ext %.1f cns %.1f) | wall %s | bases %d" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms"]["extend_ms"], d["stage_ms"]["consensus_ms"], d["stage_ms"]["extend_ms_timed_passes"], d["stage_ms"]["consensus_ms_timed_passes"], {k: round(v) for k, v in d["wall_ms"].items()}, d["corrected_bases"]))
except Exception as e: print(sys.argv[1], "FAILED", e, open("$O/c15_tmp.err").read()[-300:])
PY
grep -E "nb_consensus_job" $O/c15_tmp.err | head -3 | tail -1 | cut -c1-600
grep -E "nb_map_range" $O/c15_tmp.err | tail -1 | cut -c1-600
done
