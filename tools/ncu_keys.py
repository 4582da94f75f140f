#!/usr/bin/env python3
"""Prints the metrics we look at first from an ncu report.  usage: ncu_keys.py report.ncu-rep"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct']
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
for r in rows[2:]:
    print("==", r[h.index("Kernel Name")])
    for k in KEYS:
        if k in h:
            print("  %-70s %s %s" % (k, r[h.index(k)], rows[1][h.index(k)]))
    for i, x in enumerate(h):
        if 'issue_stalled' in x and 'per_issue_active' in x and float(r[i] or 0) > 0.05:
            print("  stall %-30s %s" % (x.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''), r[i]))
