#!/bin/bash
# usage: ncu_clip_report.sh <tag> [mangled kernel substring] [top lines] — key metrics + instructions per source line of one kernel
TAG=$1; PAT=${2:-_Z15clip_win_kernelILi3ELb0EEv12ClipFlatArgs}; TOP=${3:-45}
REP=/root/repo/gpurun_out/${TAG}_prof.ncu-rep
mkdir -p /tmp/sass_$TAG && cd /tmp/sass_$TAG && rm -f *.cubin && cuobjdump -xelf all /root/repo/graphitethree_b200/libb200cvt.so >/dev/null 2>&1
nvdisasm -g -c b200cvt.sm_100a.cubin > all.sass 2>/dev/null
bash /root/repo/scripts/sass_fn.sh all.sass "$PAT" > fn.sass
KN=$(echo "$PAT" | sed -E 's/^_Z[0-9]+([a-z_0-9]+kernel).*/\1/')
ncu -i $REP --page raw --csv 2>/dev/null > raw.csv
python - <<PY
import csv
rows=list(csv.reader(open('raw.csv'))); hdr=rows[0]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate.pct','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']
for r in rows[2:]:
    if "$KN" not in r[hdr.index('Kernel Name')]: continue
    print('---', r[hdr.index('Kernel Name')][:70])
    print('  '+'  '.join('%s=%s'%(w.split('.')[0].replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active','')[-34:], r[hdr.index(w)][:9]) for w in want if w in hdr))
    break
PY
ncu -i $REP --page source --csv --kernel-name regex:$KN 2>/dev/null > src.csv
python /root/repo/scripts/ncu_lines.py src.csv fn.sass $TOP
