#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total, mean, share."""
import csv, collections, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
    k = row['Kernel Name'].split('(')[0]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("# kernel, launches, total_us, mean_us, share_pct   (cold-cache, serialised: compare shares, not absolutes)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-44s %5d %10.1f %9.1f %6.1f" % (k[:44], v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
