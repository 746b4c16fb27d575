"""Summarise an `ncu --page raw --csv` dump: the metrics quoted in profiles/ and DESIGN.md, one block per kernel launch."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum']
want += [h for h in hdr if h.startswith('smsp__warp_issue_stalled') and h.endswith('per_warp_active.pct')]
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            v = r[i]
            if w.startswith('smsp__warp_issue_stalled'):
                try:
                    if float(v) < 2: continue
                except ValueError: pass
            print(f"{w:82s} {v} {units[i]}")
