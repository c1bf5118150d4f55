#!/usr/bin/env python3
"""Key raw metrics per kernel from `ncu -i rep --page raw --csv`."""
import csv, sys, subprocess
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct']
idx = [(w, h.index(w)) for w in want if w in h]
st = [i for i, x in enumerate(h) if 'smsp__average_warps_issue_stalled' in x and 'per_issue_active' in x]
for r in rows[2:]:
    print({w.split('.')[0][-40:]: r[i][:40] for w, i in idx})
    print("   stalls:", sorted([(round(float(r[i]), 2), h[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i in st], reverse=True)[:6])
