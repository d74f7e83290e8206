#!/bin/bash
# Round 2, GPU call 6: k_extend slots-per-warp sweep.
mkdir -p gpurun_out
O=gpurun_out
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
for v in necat_b200 build/var_s4 build/var_s16 build/var_s32; do
for c in 2 3 4; do
for g in 1 3; do
echo "== $v CTAS=$c GROUPS=$g: $(NB_EXT_CTAS=$c NB_CNS_GROUPS=$g NB_TRACE=1 $v/bin/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c.fa $W/r.fa 2>&1 | grep -E "nb_consensus" | sed 's/.*extension phase wall/ext wall/' | cut -c1-200) $(md5sum < $W/c.fa | cut -c1-8)"
done; done; done
NB_CNS_GROUPS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend' -c 1 -o $O/c6_ext -f \
    $B/oc2cns $CN -t 16 $W/vol $W/cand.bin $W/c3.fa $W/r3.fa > $O/c6_ncu.log 2>&1
tail -2 $O/c6_ncu.log
timeout 300 python -m pytest tests/test_gpu_extend.py -x -q 2>&1 | tail -2
