#!/usr/bin/env python
"""Print the kernels of the last full step found in an ncu launch list (gpu__time_duration.sum CSV)."""
import csv
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = list(csv.DictReader(lines))
idx = [i for i, x in enumerate(rows) if 'rows_zero' in x['Kernel Name'] or 'sparse_adam' in x['Kernel Name']]
a, b = idx[-2] + 1, idx[-1] + 1
tot = 0.0
for x in rows[a:b]:
    t = float(x['Metric Value']) / 1000
    tot += t
    print(f"{x['Kernel Name'][:90]:90s} {x['Grid Size']:14s} {x['Block Size']:12s} {t:8.1f}")
print(f'total {tot:.1f} us over {b - a} launches')
