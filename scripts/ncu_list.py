"""Prints the per-kernel metrics of an `ncu --csv --log-file` launch list (last N launches)."""
import csv, collections, re, sys
path = sys.argv[1]; last = int(sys.argv[2]) if len(sys.argv) > 2 else 30
lines = [l for l in open(path) if not l.startswith('==')]
byid = collections.OrderedDict()
for row in csv.DictReader(lines):
    byid.setdefault(row['ID'], {'name': re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')[:40]})[row['Metric Name']] = row['Metric Value']
short = {'gpu__time_duration.sum': 'ns', 'smsp__thread_inst_executed_per_inst_executed.ratio': 'lanes', 'smsp__inst_executed.sum': 'winst',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue%',
         'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active': 'fp64%'}
for i in list(byid)[-last:]:
    d = byid[i]
    print("%4s %-40s %s" % (i, d['name'], ' '.join("%s=%s" % (short.get(k, k[-24:]), v) for k, v in d.items() if k != 'name')))
