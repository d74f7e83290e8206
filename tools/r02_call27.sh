#!/bin/bash
# Round 2, GPU call 27: index with a coalesced table sweep (parity, launch list); link-kernel grid knob on the consensus job.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_index.py tests/test_gpu_pmov.py -x -q 2>&1 | tail -2
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
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/c27_launches_pmov.csv \
    $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand2.bin > /dev/null 2>&1
python tools/launch_summary.py $O/c27_launches_pmov.csv 2>/dev/null | tail -16 | tee $O/c27_launch_summary_pmov.txt
for c in 6 7 8 9 6 8; do
NB_CNS_LINKS_CTAS=$c timeout 300 python tools/bench_shard.py 1 2>&1 | grep shard
done
