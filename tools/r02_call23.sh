#!/bin/bash
# Round 2, GPU call 23: the queue of the longest candidates in k_extend, at N = 1 and as one rank of 2 / 4 / 8 sees it.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_extend.py "tests/test_gpu_cns.py::test_cns_kernel_variants" -x -q 2>&1 | tail -2
NB_EXT_LONGQ=0 timeout 300 python tools/bench_shard.py 1 2 4 8 2>&1 | grep shard | tee $O/c23_shard_q0.txt
NB_EXT_LONGQ=1 timeout 300 python tools/bench_shard.py 1 2 4 8 2>&1 | grep shard | tee $O/c23_shard_q1.txt
