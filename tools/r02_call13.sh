#!/bin/bash
# Round 2, GPU call 13: window k_cns_dp (lane = link of a level, no plan pass): parity, launch list, bench.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cns.py -x -q > $O/c13_pytest.log 2>&1
tail -2 $O/c13_pytest.log
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
CN="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
timeout 60 $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin > /dev/null 2>&1; echo "pmov rc=$?"
timeout 60 $B/oc2pcan -t 16 $W/vol $W/cand.bin > /dev/null 2>&1
NB_TRACE=1 timeout 90 $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa > /dev/null 2> $O/c13_exe_trace.txt; echo "cns rc=$?"
md5sum $W/c.fa $W/r.fa | cut -c1-32 | tr '\n' ' '; echo
echo "expected ecff682e6bd7218f86524b56b5545ad4 5e655a5a16d9ddc5fc9ef91eb8d9c8c3"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c13_launches_cns.csv \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c2.fa $W/r2.fa > /dev/null 2>&1
python - <<PY
import csv
rows = list(csv.reader(open("$O/c13_launches_cns.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
for r in rows[h + 1:]:
    print(r[4][:44], r[6], r[8], "%.2f ms" % (float(r[-1]) / 1e6))
PY
NB_CNS_LONG_PCT=0 NB_TRACE=1 timeout 90 $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa 2>&1 > /dev/null | grep nb_consensus_job | cut -c1-400
NB_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/c13_bench.json 2> $O/c13_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c13_bench.json").read().strip().splitlines()[-1])
    print("bench", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, d["e2e"]["ms_per_step"])
except Exception as e: print("bench failed", e)
PY
grep -E "nb_consensus_job" $O/c13_bench.err | tail -2 | cut -c1-600
