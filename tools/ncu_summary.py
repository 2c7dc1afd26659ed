#!/usr/bin/env python
"""Per-launch summary of an .ncu-rep (`ncu --set full`): the columns the DESIGN / bench roofline cite.

    tools/ncu_summary.py <report.ncu-rep> <out.csv>

Reads the report with `ncu -i <rep> --page raw --csv` and keeps duration, DRAM bytes, launch shape,
registers, shared memory, occupancy / issue utilisation and the main stall ratios."""
import csv
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active']

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
print(f'{len(rows) - 2} launches -> {out}')
