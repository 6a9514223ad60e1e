"""Print the headline raw metrics of an .ncu-rep (ncu -i ... --page raw --csv piped on stdin)."""
import csv
import sys
rows = list(csv.reader(sys.stdin))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__block_size', 'launch__grid_size', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__waves_per_multiprocessor', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__shared_mem_per_block_dynamic', 'sm__maximum_warps_per_active_cycle_pct', 'launch__occupancy_per_block_size']
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(w, [r[i] for r in rows[1:]])
