#!/bin/bash
# Round 2, GPU call 17: ncu --set full captures, kept small (one report per kernel group, summaries made on the box),
# with the branch-free walk-back of k_extend; parity of the extension tests; one bench line.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_extend.py tests/test_gpu_index.py -x -q > $O/c17_pytest.log 2>&1
tail -2 $O/c17_pytest.log
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
NB_TRACE=1 timeout 60 $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin 2>&1 > /dev/null | grep -E "phase" | tr '\n' ';'; echo
timeout 60 $B/oc2pcan -t 16 $W/vol $W/cand.bin > /dev/null 2>&1
NB_TRACE=1 timeout 90 $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa 2>&1 > /dev/null | grep -E "phase" | tr '\n' ';'; echo
md5sum $W/c.fa $W/r.fa | cut -c1-32 | tr '\n' ' '; echo
echo "expected ecff682e6bd7218f86524b56b5545ad4 5e655a5a16d9ddc5fc9ef91eb8d9c8c3"
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:'k_part1_scatter|k_part2_scatter|k_bucket_count|k_bucket_fill|k_probe|k_build_blocks|k_candidates' -c 7 -o $O/c17_pmov -f \
    $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand2.bin > $O/c17_ncu_pmov.log 2>&1; tail -1 $O/c17_ncu_pmov.log
timeout 300 $NCU -k regex:'k_extend' -c 1 -o $O/c17_ext -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c3.fa $W/r3.fa > $O/c17_ncu_ext.log 2>&1; tail -1 $O/c17_ncu_ext.log
# consensus kernels: the launches of the second slice (the first one of the "rest": largest link kernel, k_cns_dp<8>) and the chain of slice 0
timeout 400 $NCU -k regex:'k_cns_links|k_cns_dp<|k_cns_walk' --launch-skip 0 -c 6 -o $O/c17_cns -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c4.fa $W/r4.fa > $O/c17_ncu_cns.log 2>&1; tail -1 $O/c17_ncu_cns.log
for r in pmov ext cns; do python tools/ncu_summary.py $O/c17_$r.ncu-rep > $O/c17_ncu_$r.txt 2>&1; done
ls -la $O/c17_*
NB_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 > $O/c17_bench.json 2> $O/c17_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c17_bench.json").read().strip().splitlines()[-1])
    print("bench", {k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms")}, d["e2e"]["ms_per_step"], d.get("parity"), d.get("e2e_exe"), d["roofline"])
except Exception as e: print("bench failed", e)
PY
du -sh $O
