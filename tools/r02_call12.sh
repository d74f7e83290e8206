#!/bin/bash
# Round 2, GPU call 12: vectorised replay loops, previous k_cns_dp restored: parity, bench trace; k_cns_links tile variants.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_cns.py -x -q -k "matches_oracle or multi_volume" > $O/c12_pytest.log 2>&1
tail -2 $O/c12_pytest.log
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
for V in necat_b200 build/var_t128 build/var_t32; do
NB_CNS_GROUPS=1 NB_TRACE=1 timeout 90 $V/bin/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa > /dev/null 2> $O/c12_trace_$(basename $V).txt; echo "$V cns rc=$?"
grep -E "nb_consensus_job" $O/c12_trace_$(basename $V).txt | tail -1 | cut -c1-500
md5sum $W/c.fa $W/r.fa | cut -c1-32 | tr '\n' ' '; echo
done
echo "expected ecff682e6bd7218f86524b56b5545ad4 5e655a5a16d9ddc5fc9ef91eb8d9c8c3"
for g in 1 3; do
NB_TRACE=1 NB_CNS_GROUPS=$g timeout 300 python bench.py --no-cpu-baseline --steps 3 --warmup 3 > $O/c12_bench_g$g.json 2> $O/c12_bench_g$g.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c12_bench_g$g.json").read().strip().splitlines()[-1])
    print("groups $g", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, d["e2e"]["ms_per_step"])
except Exception as e: print("bench g$g failed", e)
PY
grep -E "nb_consensus_job" $O/c12_bench_g$g.err | tail -2 | cut -c1-600
done
