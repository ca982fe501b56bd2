"""Summarise `ncu -i X.ncu-rep --page raw --csv` exports into the small JSON kept under profiles/.
Usage: ncu_summary.py out.json raw1.csv [raw2.csv ...]"""
import csv, json, sys
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
out = []
for path in sys.argv[2:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        rec = {k: d.get(k) for k in KEYS}
        rec['units'] = {k: units[hdr.index(k)] for k in KEYS if k in hdr}
        rec['source'] = path.split('/')[-1]
        out.append(rec)
json.dump(out, open(sys.argv[1], 'w'), indent=1)
for rec in out:
    print(rec['Kernel Name'], rec['gpu__time_duration.sum'], 'us', rec['dram__bytes_read.sum'], '+', rec['dram__bytes_write.sum'], 'MB', rec['smsp__inst_executed.sum'], 'inst')
