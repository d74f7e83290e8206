#!/bin/bash
# Round 2, GPU call 16: ncu --set full captures of the kernels as they are now (extension, consensus, index, seed probe + vote),
# and the launch list of one bench step.
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
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_part|k_bucket|k_probe|k_build_blocks|k_candidates' -c 12 -o $O/c16_pmov -f \
    $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand2.bin > $O/c16_ncu_pmov.log 2>&1; tail -1 $O/c16_ncu_pmov.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -c 1 -o $O/c16_ext -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c3.fa $W/r3.fa > $O/c16_ncu_ext.log 2>&1; tail -1 $O/c16_ncu_ext.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_cns_links|k_cns_dp|k_cns_walk|k_cns_prep|k_cns_ckpt' -c 24 -o $O/c16_cns -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c4.fa $W/r4.fa > $O/c16_ncu_cns.log 2>&1; tail -1 $O/c16_ncu_cns.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c16_launches_bench.csv \
    python bench.py --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
python tools/launch_summary.py $O/c16_launches_bench.csv 2>/dev/null | tail -32 | tee $O/c16_launch_summary_bench.txt
ls -la $O/c16_*.ncu-rep
