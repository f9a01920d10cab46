#!/usr/bin/env python
"""Key raw metrics of every kernel in an .ncu-rep.  usage: ncu_summary.py <rep>"""
import csv, re, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
pat = re.compile(r'^(gpu__time_duration.sum|smsp__inst_executed.sum|sm__inst_executed.avg.per_cycle_active|sm__warps_active.avg.pct_of_peak_sustained_active|smsp__issue_active.avg.pct_of_peak_sustained_active|launch__registers_per_thread|launch__grid_size|launch__block_size|l1tex__t_sector_hit_rate.pct|lts__t_sector_hit_rate.pct|dram__bytes_read.sum|dram__bytes_write.sum|smsp__thread_inst_executed_per_inst_executed.ratio|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum|smsp__average_warps_issue_stalled_.*_per_issue_active.ratio|sm__inst_executed_pipe_(alu|lsu|fma|xu).avg.pct_of_peak_sustained_active|sm__inst_executed_pipe_.*pct_of_peak_sustained_active)$')
for r in rows[2:]:
    print(r[h.index('Kernel Name')][:90])
    for i, w in enumerate(h):
        if pat.match(w) and r[i] not in ('0', '', 'n/a'):
            try:
                v = float(r[i].replace(',', ''))
            except ValueError:
                continue
            if 'stalled' in w and v < 0.05:
                continue
            if 'pipe' in w and v < 1:
                continue
            print(f"   {w} [{u[i]}] = {r[i]}")
