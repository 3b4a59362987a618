#!/usr/bin/env python
"""Print the handful of `ncu --set full` metrics the profiles/ summaries quote, one block per launch.
    python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "lts__t_sector_hit_rate.pct"]
ki = h.index("Kernel Name")
for r in rows[2:]:
    print(r[ki][:110])
    for k in want:
        if k in h:
            print(f"    {k:75s} {r[h.index(k)]:>16s} {rows[1][h.index(k)]}")
    stalls = [(float(r[i]), k) for i, k in enumerate(h) if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and r[i]]
    print("    stalls per issued instruction:", ", ".join(f"{k.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, k in sorted(stalls, reverse=True)[:6]))
