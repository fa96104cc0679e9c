"""summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one training step (between two adam_kernel
launches), per-kernel totals and the launch sequence.   python scripts/summarise_launches.py gpurun_out/launches.csv [-v]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(x['Kernel Name'], float(x['Metric Value'].replace(',', '')) / 1e3, x['Grid Size'], x['Stream'])
        for x in csv.DictReader(lines)]
adam = [i for i, r in enumerate(rows) if 'adam' in r[0]]
step = rows[adam[0] + 1:adam[1] + 1] if len(adam) >= 2 else rows
tot = sum(t for _, t, _, _ in step)
print('step launches %d total %.1f us' % (len(step), tot))


def short(n):
    return re.sub(r'\(.*', '', n).replace('<unnamed>::', '').replace('void ', '')


agg = collections.OrderedDict()
for n, t, g, s in step:
    a = agg.setdefault(short(n), [0, 0.])
    a[0] += 1
    a[1] += t
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-52s %3d %9.1f us %5.1f%%' % (k[:52], c, t, 100 * t / tot))
if '-v' in sys.argv:
    print()
    for n, t, g, s in step:
        print('%-48s %8.1f %s s%s' % (short(n)[:48], t, g, s))
