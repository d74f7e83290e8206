#!/bin/bash
# Round 2, GPU call 32: the read set of BASELINE config 4 at full size on ONE GPU (140 Mb x 40 = 5.6 Gbp: three volumes, six
# volume pairs, eight partitions; the merged read store of oc2cns is beyond 2^32 bases) through the five executables; the
# reference's oc2cns on the same volumes and partition files for one whole partition whose reads lie beyond 2^32 bases.
mkdir -p gpurun_out
O=gpurun_out
df -h /tmp | tail -1
timeout 1100 python tools/scale_volumes.py --genome 140000000 --depth 40 --ref-cns-only --ref-parts=-2 --out $O/c32_scale_config4.json 2>&1 | tail -30 | cut -c1-1200
