#!/bin/bash
# Round 2, GPU call 26: the index with vector sweeps (parity + launch list), consensus link grid knob, bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_index.py tests/test_gpu_pmov.py -x -q 2>&1 | tail -2
W=/tmp/nbw; mkdir -p $W
timeout 120 python - <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import gen_reads
from necat_b200 import formats
_, reads = gen_reads.generate(4600000, 40, 20260925)
formats.write_workdir("$W/vol", reads)
PY
B=necat_b200/bin
PM="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c26_launches_pmov.csv \
    $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand2.bin > /dev/null 2>&1
python tools/launch_summary.py $O/c26_launches_pmov.csv 2>/dev/null | tail -16 | tee $O/c26_launch_summary_pmov.txt
md5sum $W/cand2.bin
for env in "A=1" "NB_CNS_LINKS_CTAS=8" "NB_CNS_LINKS_CTAS=10" "NB_CNS_LONG_PCT=5"; do
env $env timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 2 > $O/c26_tmp.json 2> $O/c26_tmp.err
python - "$env" <<PY
import json, sys
try:
    d = json.loads(open("$O/c26_tmp.json").read().strip().splitlines()[-1])
    print(sys.argv[1], "| ms/step %.1f e2e %.1f | idx %.1f map %.1f ext %.1f cns %.1f | wall %s" % (d["ms_per_step"], d["e2e"]["ms_per_step"], d["stage_ms"]["index_ms"], d["stage_ms"]["map_ms"], d["stage_ms"]["extend_ms"], d["stage_ms"]["consensus_ms"], {k: round(v) for k, v in d["wall_ms"].items()}))
except Exception as e: print(sys.argv[1], "FAILED", e, open("$O/c26_tmp.err").read()[-300:])
PY
done
