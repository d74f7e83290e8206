#!/bin/bash
# Round 2, GPU call 29: mapping tests with the ultra-long reads (24-bit probe numbers, unstaged probe path, templates of a
# million positions through oc2cns); slice / long-share knobs of the consensus launch plan on the bench workload.
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_pmov.py tests/test_gpu_index.py -x -q 2>&1 | tail -15 | tee $O/c29_pytest.log
for env in "A=1" "NB_CNS_SLICES=4" "NB_CNS_SLICES=5" "NB_CNS_LONG_PCT=12" "NB_CNS_LONG_PCT=12 NB_CNS_SLICES=4" "NB_CNS_LONG_PCT=5 NB_CNS_SLICES=4" "A=2"; do
env $env timeout 300 python tools/bench_shard.py 1 2>&1 | grep shard | tee -a $O/c29_knobs.txt
done
