"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv --log-file` launch list by kernel name.
usage: ncu_agg.py <launches.csv> [skip_first_n]"""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum':
        continue
    n += 1
    if n <= skip:
        continue
    name = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '')[:60]
    v = float(row['Metric Value'].replace(',', ''))
    unit = row['Metric Unit']
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v if unit in ('us', 'usecond') else v * 1e3 if unit in ('ms', 'msecond') else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("launches %d  total %.1f us" % (n - skip, tot))
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%8.1f us %5.1f%%  n=%4d  avg %8.2f us  %s" % (t, 100 * t / tot, c, t / c, name))
