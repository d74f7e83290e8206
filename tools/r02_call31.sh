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
