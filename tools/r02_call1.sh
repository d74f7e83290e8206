#!/bin/bash
# Round 2, GPU call 1 (round-1 code): evidence the verdict asked for.
#   1. bench line at N = 1 on this pod (baseline for the round);
#   2. ncu --set full of the index + seed-vote kernels at E. coli size (oc2pmov);
#   3. two chained correction iterations at E. coli size against the reference binaries (iteration 2: FAST + -r 1);
#   4. launch list + ncu --set full of the -r 1 rescue kernels on the iteration-2 input.
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python bench.py --no-cpu-baseline > $O/c1_bench_n1.json 2> $O/c1_bench_n1.err
tail -c 600 $O/c1_bench_n1.json; echo
W=/tmp/nbw; mkdir -p $W
python - <<PY
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import gen_reads
from necat_b200 import formats
_, reads = gen_reads.generate(4600000, 40, 20260925)
formats.write_workdir("$W/vol", reads)
PY
B=necat_b200/bin
PM="-k 15 -z 10 -q 500 -b 2000 -s 3 -n 500 -a 1000 -d 0.25 -e 0.5 -m 500 -j 0 -u 1 -i 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_probe|k_build_blocks|k_candidates|k_scan_kmers|k_scan_apply|k_mask_reduce|k_sort_short_runs|k_sort_long_runs' -c 16 \
    -o $O/c1_pmov -f $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand.bin > $O/c1_ncu_pmov.log 2>&1
tail -3 $O/c1_ncu_pmov.log
timeout 900 python tools/full_parity.py --iterations 2 --out $O/c1_parity_iter2.json --keep $W/it > $O/c1_parity_iter2.log 2>&1
grep -E '^\{"iteration"|^GPU|failed' $O/c1_parity_iter2.log | cut -c1-500
CN2="-a 2000 -x 4 -y 12 -l 1000 -e 0.5 -p 0.8 -u 0 -r 1 -f 0"
ls $W/it | head -40 > $O/c1_it_ls.txt
NB_TRACE=1 $B/oc2cns $CN2 -t 16 $W/it/gpu_it2_wrk $W/it/gpu_it2_cand.bin.part $W/c.fa $W/r.fa > $O/c1_cns_r1_trace.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/c1_launches_cns_r1.csv \
    $B/oc2cns $CN2 -t 16 $W/it/gpu_it2_wrk $W/it/gpu_it2_cand.bin.part $W/c.fa $W/r.fa > $O/c1_ncu_cns_r1_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_rescue' -c 9 -o $O/c1_rescue -f \
    $B/oc2cns $CN2 -t 16 $W/it/gpu_it2_wrk $W/it/gpu_it2_cand.bin.part $W/c.fa $W/r.fa > $O/c1_ncu_rescue.log 2>&1
tail -3 $O/c1_ncu_rescue.log
ls -la $O | tail -20
