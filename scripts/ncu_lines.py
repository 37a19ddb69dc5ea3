"""Joins an `ncu --page source --csv` SASS listing with nvdisasm -g line info: instructions executed per source line.
usage: ncu_lines.py <source.csv> <function.sass (nvdisasm -g -c excerpt)> [top]"""
import csv, re, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
# several launches may be listed one after the other: keep the first
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
if len(starts) > 1: rows = rows[starts[0]:starts[1]]
hdr = rows[1]; ix = {h: j for j, h in enumerate(hdr)}
sass = []
cur = None
for line in open(sys.argv[2]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.search(r'/\*([0-9a-f]{4,})\*/\s+(.*?);', line)
    if m:
        sass.append((int(m.group(1), 16), m.group(2).strip(), cur))
data = rows[2:]
assert len(data) == len(sass), (len(data), len(sass))
agg = collections.defaultdict(lambda: [0, 0, 0])
tot = 0
for r, (off, txt, loc) in zip(data, sass):
    ie = int(r[ix['Instructions Executed']]); te = int(r[ix['Thread Instructions Executed']])
    agg[loc][0] += ie; agg[loc][1] += te; agg[loc][2] += 1
    tot += ie
src = {}
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
print("total warp instructions", tot)
for loc, (ie, te, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    f, l = loc if loc else ("?", 0)
    if f not in src:
        try: src[f] = open('/root/repo/graphitethree_b200/csrc/' + f).read().split('\n')
        except Exception: src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ''
    print("%10d %5.1f%% lanes=%5.1f sass=%3d %s:%d | %s" % (ie, 100.0 * ie / tot, te / max(ie, 1), n, f, l, text))
