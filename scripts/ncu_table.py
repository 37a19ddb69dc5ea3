"""Markdown table of the metrics that matter from `ncu -i <rep> --page raw --csv` (one column per captured launch).
usage: ncu -i x.ncu-rep --page raw --csv > raw.csv ; python scripts/ncu_table.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
ix = {h: i for i, h in enumerate(hdr)}
names = [r[ix['Kernel Name']].replace('void ', '').split('(')[0][:28] for r in data]
print('| metric | ' + ' | '.join(names) + ' |')
print('|---|' + '---|' * len(names))
for w in want:
    if w in ix:
        short = w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
        print('| %s (%s) | ' % (short, units[ix[w]]) + ' | '.join(r[ix[w]][:10] for r in data) + ' |')
