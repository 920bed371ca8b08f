"""Print selected metrics of an `ncu --page raw --csv` export (developer tool).  python scripts/ncu_raw.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_active.avg', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__grid_size', 'launch__block_size']
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print('---', d.get('Kernel Name', '')[:60])
    for h, u in zip(hdr, units):
        if h in want or ('issue_stalled' in h and 'per_issue_active' in h and float(d[h] or 0) > 0.3):
            print('  %-80s %-12s %s' % (h, u, d[h]))
