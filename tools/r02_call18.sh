#!/bin/bash
# Round 2, GPU call 18: k_extend occupancy variants; iteration 2 (-r 1) timed with its launch list; bench with the executables
# stand-alone and behind the resident helper.
mkdir -p gpurun_out
O=gpurun_out
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
for V in necat_b200 build/var_s8c4 build/var_s16c4 necat_b200; do
NB_TRACE=1 timeout 90 $V/bin/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa 2>&1 > /dev/null | grep -E "nb_consensus_job" | sed 's/.*gpu ms/gpu ms/' | tr '\n' ' '; echo " <- $V"
md5sum $W/c.fa | cut -c1-32
done
echo "expected ecff682e6bd7218f86524b56b5545ad4"
# ---- iteration 2 (FAST overlap options, -r 1) on the corrected reads of iteration 1, this repo's executables only
timeout 300 python tools/full_parity.py --iterations 2 --skip-ref --keep /tmp/it2 --out $O/c18_iter2.json > $O/c18_iter2.log 2>&1; tail -2 $O/c18_iter2.log | cut -c1-300
I2=/tmp/it2
ls $I2 | head -30
CN2="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"
NB_TRACE=1 timeout 120 $B/oc2cns $CN2 -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c2.fa $W/r2.fa > /dev/null 2> $O/c18_r1_trace.txt; echo "cns -r 1 rc=$?"
grep -E "phase|nb_consensus_job|rescue" $O/c18_r1_trace.txt | cut -c1-500 | tail -12
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/c18_launches_r1.csv \
    $B/oc2cns $CN2 -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c3.fa $W/r3.fa > /dev/null 2>&1
python tools/launch_summary.py $O/c18_launches_r1.csv 2>/dev/null | tail -20 | tee $O/c18_launch_summary_r1.txt
# ---- the bench line, with the reference binaries beside it (parity, e2e_exe, resident helper)
NB_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 > $O/c18_bench.json 2> $O/c18_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c18_bench.json").read().strip().splitlines()[-1])
    print("bench", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, d["e2e"]["ms_per_step"], d.get("parity"))
    print("e2e_exe", json.dumps(d.get("e2e_exe"))[:1500])
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e: print("bench failed", e)
PY
pgrep -c nb_helperd
