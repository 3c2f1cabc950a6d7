#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) into the handful of metrics the design notes quote."""
import csv, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fma.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores']
text = open(sys.argv[1]).read() if sys.argv[1].endswith('.csv') else subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(text.splitlines()))
hdr = rows[0]
for i, h in enumerate(hdr):
    if h in KEYS or h == 'Kernel Name' or ('issue_stalled' in h and 'per_issue_active' in h and 'pcsamp' not in h):
        vals = [r[i] for r in rows[1:]]
        print(f"{h:90s} {vals[0]:>12s} " + "  ".join(f"{v[:44]:>14s}" for v in vals[1:]))
