#!/usr/bin/env python
"""Prints the headline counters of an .ncu-rep (read here with `ncu -i`, no GPU needed). Development helper."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__average_warp', 'smsp__warp_issue_stalled', 'sm__cycles_elapsed.avg', 'launch__shared_mem_per_block_dynamic',
        'lts__t_bytes.sum', 'sm__cycles_active.avg', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')])
    for i, h in enumerate(hdr):
        if any(h.startswith(w) for w in WANT):
            if 'warp_issue_stalled' in h and not h.endswith('_per_warp_active.pct'):
                continue
            print(f'  {h} [{units[i]}] = {r[i]}')
