#!/bin/bash
# Round 2, GPU call 7: DP v4 + k_extend s16: parity tests, bench (groups 1 and 3), launch list, ncu of k_cns_dp/plan.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_cns.py tests/test_gpu_extend.py -x -q > $O/c7_pytest.log 2>&1
tail -3 $O/c7_pytest.log
for g in 3 1; do
NB_CNS_GROUPS=$g timeout 400 python bench.py --no-cpu-baseline > $O/c7_bench_g$g.json 2> $O/c7_bench_g$g.err
python - <<PY
import json
d = json.load(open("$O/c7_bench_g$g.json"))
print("groups $g", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "corrected_bases", "candidates")})
PY
done
W=/tmp/nbw; mkdir -p $W
python - <<PY
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
$B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin > /dev/null 2>&1
$B/oc2pcan -t 16 $W/vol $W/cand.bin > /dev/null 2>&1
$B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa > /dev/null 2>&1
md5sum $W/c.fa $W/r.fa | tee $O/c7_md5.txt
echo "expected ecff682e6bd7218f86524b56b5545ad4 5e655a5a16d9ddc5fc9ef91eb8d9c8c3"
NB_CNS_GROUPS=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c7_launches_cns.csv \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c2.fa $W/r2.fa > /dev/null 2>&1
python tools/launch_summary.py $O/c7_launches_cns.csv 2>/dev/null | tail -13
NB_CNS_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cns_dp|k_cns_plan' -c 3 -o $O/c7_dp -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c3.fa $W/r3.fa > $O/c7_ncu.log 2>&1
tail -2 $O/c7_ncu.log
