#!/bin/bash
# Round 2, GPU call 9: overlap mode (per-group consensus behind the next group's extension) vs one group.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cns.py -x -q > $O/c9_pytest.log 2>&1
tail -3 $O/c9_pytest.log
run() {
  env "$@" timeout 400 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > $O/c9_tmp.json 2> $O/c9_tmp.err
  python - "$*" <<PY
import json, sys
try:
    d = json.loads(open("$O/c9_tmp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms/step %.1f e2e %.1f | ext %.1f cns %.1f | wall %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms"]["extend_ms_timed_passes"], d["stage_ms"]["consensus_ms_timed_passes"], {k: round(v) for k, v in d["wall_ms"].items()}))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("$O/c9_tmp.err").read()[-300:])
PY
}
run NB_CNS_GROUPS=1
run NB_CNS_GROUPS=2
run NB_CNS_GROUPS=3
run NB_CNS_GROUPS=4
run NB_CNS_GROUPS=3 NB_CNS_OVERLAP=0
run NB_CNS_GROUPS=3 NB_EXT_CTAS=3
cp $O/c9_tmp.json $O/c9_bench_last.json
