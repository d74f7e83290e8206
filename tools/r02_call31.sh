#!/bin/bash
# Round 2, GPU call 31: the whole GPU suite, the bench line and the ncu launch list of the bench command on the final code.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c31_pytest_gpu.log 2>&1
tail -3 $O/c31_pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
NB_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 3 > $O/c31_bench.json 2> $O/c31_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$O/c31_bench.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("ms_per_step", "stage_ms", "wall_ms", "parity", "gpu_launches", "clocks")})
    print("e2e", d["e2e"], "roofline", {k: d["roofline"][k] for k in ("achieved", "frac", "traffic")})
    print("e2e_exe", d.get("e2e_exe", {}).get("ms_per_step"), d.get("e2e_exe", {}).get("resident_helper", {}).get("later_passes", {}).get("ms_per_step"))
    print("cpu_baseline", d.get("cpu_baseline"))
except Exception as e: print("bench failed", e, open("$O/c31_bench.err").read()[-600:])
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $O/c31_launches_bench_ncu.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/launch_summary.py $O/c31_launches_bench_ncu.csv 2>/dev/null | tail -40 | tee $O/c31_launch_summary_bench.txt
# ---- ncu --set full of the index and mapping kernels as they are now (the bucket kernels changed after call 17)
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
NCU="ncu --set full --clock-control none --import-source on"
timeout 300 $NCU -k regex:'k_part1_scatter|k_part2_scatter|k_bucket_count|k_bucket_fill|k_probe|k_build_blocks|k_candidates' -c 7 -o $O/c31_pmov -f \
    $B/oc2pmov $PM -t 16 $W/vol 0 $W/cand2.bin > $O/c31_ncu_pmov.log 2>&1; tail -1 $O/c31_ncu_pmov.log
python tools/ncu_summary.py $O/c31_pmov.ncu-rep > $O/c31_ncu_pmov.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes" $O/c31_ncu_pmov.txt | head -40
rm -f $O/c31_pmov.ncu-rep
