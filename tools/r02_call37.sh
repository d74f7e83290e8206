#!/bin/bash
# Round 2, GPU call 37: L2 fetch granularity (cudaLimitMaxL2FetchGranularity 32 / 64 / 128 B) against the seed probe's sector
# traffic: phases of oc2pmov and GPU ms of the consensus job on the bench workload.
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
for g in 0 32 64 128 0 32; do
echo "== NB_L2_FETCH=$g" | tee -a $O/c37_l2_fetch.txt
NB_L2_FETCH=$g NB_TRACE=1 timeout 60 $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin 2>&1 > /dev/null | grep -E "L2 fetch|nb_map_range|phase\] index" | cut -c1-200 | tee -a $O/c37_l2_fetch.txt
if [ ! -f $W/cand.bin.partitions ]; then cp $W/cand.bin $W/cp.bin; timeout 60 $B/oc2pcan -t 16 $W/vol $W/cp.bin > /dev/null 2>&1; fi
NB_L2_FETCH=$g NB_TRACE=1 timeout 90 $B/oc2cns $CN -t 16 $W/vol $W/cp.bin $W/c.fa $W/r.fa 2>&1 > /dev/null | grep -E "nb_consensus_job" | sed 's/.*gpu ms/gpu ms/' | tee -a $O/c37_l2_fetch.txt
md5sum $W/c.fa | cut -c1-32
done
