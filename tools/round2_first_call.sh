#!/bin/bash
# The first GPU call of the next round, in one payload (about 8 minutes of box time):
#   1. every GPU test, including the ones written after round 1's GPU budget was spent
#      (tests/test_iteration_flow.py::test_two_iterations_gpu, tests/test_zz_gpu_edges.py);
#   2. the bench line at N = 1;
#   3. TWO chained correction iterations at E. coli size, this repo's five executables against the reference's five,
#      iteration 2 with the FAST overlap options and oc2cns -r 1 (the first run of -r 1 at scale): parity + wall clock
#      per stage (DESIGN.md "Next", item 1).
# usage:  gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -c 700 gpurun_out/bench_n1.json; echo
timeout 900 python tools/full_parity.py --iterations 2 --out gpurun_out/parity_iter2.json > gpurun_out/parity_iter2.log 2>&1
grep -E '^\{"iteration"|^GPU|failed' gpurun_out/parity_iter2.log | cut -c1-400
# random options / read sets, this repo's executables against the oracle (each case ~5 s)
timeout 600 python tools/soak_oracle.py 1 60 gpu > gpurun_out/soak_gpu.log 2>&1
grep -v ' OK ' gpurun_out/soak_gpu.log | cut -c1-300 | tail -8
