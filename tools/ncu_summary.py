#!/usr/bin/env python3
"""Condenses an `ncu --set full` report into the per-kernel numbers DESIGN.md and bench.py quote
(duration, DRAM bytes, IPC, occupancy, hit rates) plus the hottest source lines of each kernel.
Usage: ncu_summary.py report.ncu-rep > profiles/<name>.txt"""
import csv
import os
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
ki = h.index("Kernel Name")
seen = {}
print(f"# {os.path.basename(rep)}: ncu --set full --clock-control none --import-source on (per launch; cold caches, serialised)")
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace(" ", "")          # template instances keep their arguments: k_cns_dp<8,128>
    n = seen.get(name, 0)
    seen[name] = n + 1
    print(f"\n== {name} (launch #{n})")
    for w in want:
        if w in h:
            i = h.index(w)
            print(f"  {w:62s} {r[i]:>18s} {units[i]}")
here = os.path.dirname(os.path.abspath(__file__))
for name in seen:
    print(f"\n== hottest source lines of {name} (launch #0): share of warp instructions / of stall samples")
    out = subprocess.run([sys.executable, os.path.join(here, "ncu_lines.py"), rep, name.split("<")[0], "0", "14"], stdout=subprocess.PIPE, text=True).stdout
    print("  " + out.replace("\n", "\n  ").rstrip())
