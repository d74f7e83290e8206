#!/bin/bash
# Round 2, GPU call 35: the bench line twice more without the reference legs (are the 65 ms stalls of two timed passes of call 34 the box or the build?)
mkdir -p gpurun_out
O=gpurun_out
for i in 1 2; do
NB_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/c35_bench_$i.json 2> $O/c35_bench_$i.err
python - <<PY
import json
d = json.loads(open("$O/c35_bench_$i.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, "e2e", d["e2e"]["ms_per_step"])
PY
grep nb_consensus_job $O/c35_bench_$i.err | sed 's/ | summed.*//' | cut -c1-120
done
