#!/bin/bash
# Round 2, GPU call 19: iteration 2 (-r 1) timed with its launch list and an ncu capture of the rescue kernels; ncu capture of k_cns_dp.
mkdir -p gpurun_out
O=gpurun_out
B=necat_b200/bin
W=/tmp/nbw; mkdir -p $W
timeout 300 python tools/full_parity.py --iterations 2 --skip-ref --keep /tmp/it2 --out $O/c19_iter2.json > $O/c19_iter2.log 2>&1; tail -1 $O/c19_iter2.log | cut -c1-300
I2=/tmp/it2
OLD=$(head -1 $I2/gpu_it2_wrk/volume_names.txt | sed 's|/gpu_it2_wrk/.*||')
sed -i "s|$OLD|$I2|g" $I2/gpu_it2_wrk/volume_names.txt; cat $I2/gpu_it2_wrk/volume_names.txt
CN2="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"
for k in 1 2; do
NB_TRACE=1 timeout 120 $B/oc2cns $CN2 -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c2.fa $W/r2.fa > /dev/null 2> $O/c19_r1_trace.txt; echo "cns -r 1 rc=$?"
done
grep -E "phase|nb_consensus_job|rescue" $O/c19_r1_trace.txt | cut -c1-500 | tail -14
NB_TRACE=1 timeout 120 $B/oc2cns ${CN2/-r 1/-r 0} -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c2b.fa $W/r2b.fa 2>&1 > /dev/null | grep -E "nb_consensus_job" | cut -c1-500
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/c19_launches_r1.csv \
    $B/oc2cns $CN2 -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c3.fa $W/r3.fa > /dev/null 2>&1
python tools/launch_summary.py $O/c19_launches_r1.csv 2>/dev/null | tail -22 | tee $O/c19_launch_summary_r1.txt
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:'k_rescue' -c 6 -o $O/c19_rescue -f \
    $B/oc2cns $CN2 -t 16 $I2/gpu_it2_wrk $I2/gpu_it2_cand.bin.part $W/c4.fa $W/r4.fa > $O/c19_ncu_rescue.log 2>&1; tail -1 $O/c19_ncu_rescue.log
python tools/ncu_summary.py $O/c19_rescue.ncu-rep > $O/c19_ncu_rescue.txt 2>&1
# k_cns_dp on the bench workload (iteration 1)
timeout 120 python - <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import gen_reads
from necat_b200 import formats
_, reads = gen_reads.generate(4600000, 40, 20260925)
formats.write_workdir("$W/vol", reads)
PY
PM="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CN="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
timeout 60 $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin > /dev/null 2>&1
timeout 60 $B/oc2pcan -t 16 $W/vol $W/cand.bin > /dev/null 2>&1
timeout 300 $NCU -k regex:'k_cns_dp$' -c 2 -o $O/c19_dp -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c5.fa $W/r5.fa > $O/c19_ncu_dp.log 2>&1; tail -1 $O/c19_ncu_dp.log
python tools/ncu_summary.py $O/c19_dp.ncu-rep > $O/c19_ncu_dp.txt 2>&1
ls -la $O/c19_*; du -sh $O
