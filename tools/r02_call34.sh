#!/bin/bash
# Round 2, GPU call 34: the whole GPU suite, smoke and the bench line after the host-side changes (records packed in the
# library, volumes merged by whole words, context created beside the volume load); then the phases of oc2pmov / oc2cns on
# the config-4-size set again (tools/r02_call33.sh).
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c34_pytest_gpu.log 2>&1
tail -3 $O/c34_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NB_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 > $O/c34_bench.json 2> $O/c34_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c34_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "parity", "gpu_launches")})
    print("e2e", d["e2e"]["ms_per_step"], "e2e_exe", d.get("e2e_exe", {}).get("ms_per_step"), d.get("e2e_exe", {}).get("stage_s"), d.get("e2e_exe", {}).get("resident_helper", {}).get("later_passes", {}))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e: print("bench failed", e, open("$O/c34_bench.err").read()[-600:])
PY
sed -e 's/c33_/c34_/g' tools/r02_call33.sh > /tmp/c34_scale.sh; bash /tmp/c34_scale.sh
