#!/usr/bin/env python
"""Print the headline metrics of every kernel in an .ncu-rep (raw page)."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_warps', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
idx = [hdr.index(w) for w in want if w in hdr]
for r in rows[2:]:
    print('---')
    for i in idx:
        v = r[i]
        try:
            if float(v.replace(',', '')) == 0 and 'stalled' in hdr[i]:
                continue
        except ValueError:
            pass
        print("  %-88s %s %s" % (hdr[i], v[:80], rows[1][i]))
