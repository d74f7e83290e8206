#!/bin/bash
# Round 2, GPU call 38: the reference's unmodified `necat.pl correct` with this repo's executables in its bin directory on a
# B200 against the same run with the reference's binaries (tests/test_necat_pl.py); smoke on the rebuilt library.
mkdir -p gpurun_out
O=gpurun_out
true
timeout 400 python -m pytest tests/test_necat_pl.py -x -q -m gpu 2>&1 | tail -12 | tee $O/c38_necat_pl.log
