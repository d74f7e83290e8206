#!/bin/bash
# Round 2, GPU call 10: library variants and pipeline knobs on the consensus job (warm, through the C-ABI).
mkdir -p gpurun_out
O=gpurun_out
L=necat_b200/lib/libnecat_b200.so
timeout 600 python -m pytest tests/test_gpu_cns.py tests/test_gpu_extend.py -x -q > $O/c10_pytest.log 2>&1
tail -3 $O/c10_pytest.log
for g in 1 3; do
NB_CNS_GROUPS=$g NB_CNS_OVERLAP=0 python tools/bench_cns_variants.py --one $L
NB_CNS_GROUPS=$g NB_CNS_OVERLAP=0 NB_CNS_LONG_PCT=0 python tools/bench_cns_variants.py --one $L
NB_CNS_GROUPS=$g NB_CNS_OVERLAP=0 python tools/bench_cns_variants.py --one build/var_s8c4/lib/libnecat_b200.so
done
NB_CNS_GROUPS=1 NB_CNS_LONG_PCT=4 python tools/bench_cns_variants.py --one $L
NB_CNS_GROUPS=1 NB_CNS_LONG_PCT=15 python tools/bench_cns_variants.py --one $L
NB_CNS_GROUPS=2 NB_CNS_OVERLAP=0 python tools/bench_cns_variants.py --one $L
NB_TRACE=1 NB_CNS_GROUPS=1 python tools/bench_cns_variants.py --one $L 2>&1 | grep -E "nb_consensus_job|wall" | tail -3 | cut -c1-700
NB_TRACE=1 NB_CNS_GROUPS=3 NB_CNS_OVERLAP=0 python tools/bench_cns_variants.py --one $L 2>&1 | grep -E "nb_consensus_job|wall" | tail -3 | cut -c1-700
