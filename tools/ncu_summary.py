#!/usr/bin/env python
"""Print the handful of metrics we read from an `ncu --page raw --csv` export (one row per profiled launch)."""
import csv, sys
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct', 'l1tex__throughput.avg.pct',
        'lts__throughput.avg.pct', 'dram__throughput.avg.pct', 'sm__warps_active.avg.pct_of_peak', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct', 'sm__inst_executed.avg.per_cycle_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'launch__registers_per_thread', 'launch__occupancy_limit', 'launch__grid_size',
        'launch__block_size', 'smsp__average_warps_issue_stalled', 'sm__pipe_alu_cycles_active.avg.pct', 'smsp__inst_executed_pipe',
        'sm__inst_executed_pipe', 'achieved_occupancy', 'sm__maximum_warps_per_active_cycle_pct', 'smsp__thread_inst_executed_per_inst_executed']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("=====", r[4][:90], r[7], r[8])
    for h, u, v in zip(hdr, units, r):
        if any(k in h for k in KEYS) and v not in ('', '0', 'n/a'):
            if 'warps_issue_stalled' in h and '_per_issue_active' not in h: continue
            try:
                if float(v.replace(',', '')) == 0: continue
            except ValueError: pass
            print(f"{h[:100]:100s} {u:14s} {v}")
