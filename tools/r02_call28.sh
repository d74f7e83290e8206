#!/bin/bash
# Round 2, GPU call 28: the executables at the volume sizes of BASELINE configs 4-5 on one GPU (2.4 Gbp: two volumes cut at
# 2 Gbp, three partitions) against the reference binaries (tools/scale_volumes.py).
mkdir -p gpurun_out
O=gpurun_out
df -h /tmp | tail -1; grep -E "MemTotal|MemAvailable" /proc/meminfo; nproc
timeout 900 python tools/scale_volumes.py --genome 120000000 --depth 20 --out $O/c28_scale_volumes.json 2>&1 | tail -40
