#!/bin/bash
# Wall-clock of the drop-in executables and of their phases (NB_TRACE=1) on the bench workload; run on the GPU box.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
cd "$W"
python - <<PY
import sys
sys.path.insert(0, "$ROOT"); sys.path.insert(0, "$ROOT/tools")
import gen_reads
from necat_b200 import formats
_, reads = gen_reads.generate(4600000, 40, 20260925)
formats.write_workdir("vol", reads)
PY
B=$ROOT/necat_b200/bin
PM="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
CN="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 0 -f 0"
export NB_TRACE=1
for i in 1 2; do
  t0=$(date +%s.%N); $B/oc2pmov $PM -t 16 vol 0 cand.bin > pm.log 2> pm.err; t1=$(date +%s.%N)
  grep phase pm.err; echo "oc2pmov wall $(python -c "print(round($t1 - $t0, 3))") s"
done
$B/oc2pcan -t 16 vol cand.bin > /dev/null
for i in 1 2; do
  t0=$(date +%s.%N); $B/oc2cns $CN -t 16 vol cand.bin cns.fa raw.fa > cn.log 2> cn.err; t1=$(date +%s.%N)
  grep -E "phase|nb_consensus" cn.err | cut -c1-300; echo "oc2cns wall $(python -c "print(round($t1 - $t0, 3))") s"
done
rm -rf "$W"
