#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): BASELINE config 3 -- yeast-size genome, two correction iterations on 2 GPUs, against the reference chain.
mkdir -p gpurun_out
O=gpurun_out
export NB_CONFIG3_DIR=/tmp/nb_config3
# a small run first: the glue
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 tools/config3_yeast.py --genome 400000 --depth 30 --out $O/c24_config3_small.json > $O/c24_small.log 2>&1
grep -E "GPU iteration|REF iteration|Error|error" $O/c24_small.log | cut -c1-400 | tail -6
if grep -q '"cns_identical": true' $O/c24_config3_small.json 2>/dev/null; then
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29578 tools/config3_yeast.py --out $O/c24_config3_yeast_2gpu.json > $O/c24_yeast.log 2>&1
grep -E "GPU iteration|REF iteration|Error|error" $O/c24_yeast.log | cut -c1-600 | tail -6
else
tail -20 $O/c24_small.log
fi
