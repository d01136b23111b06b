"""Print the kernels of the last full training step of an ncu launch list (gpurun_out/<tag>_launches.csv)."""
import csv, sys
tag = sys.argv[1]
lines = [l for l in open(f'/root/repo/gpurun_out/{tag}_launches.csv') if not l.startswith('==')]
rows = list(csv.DictReader(lines))
seq = [(r['Kernel Name'][:72], float(r['Metric Value']) / 1000, r.get('Grid Size')) for r in rows]
first = 'presplit' if any('presplit' in s[0] for s in seq) else 'gather'
idx = [i for i, s in enumerate(seq) if first in s[0]]
a, b = idx[-2], idx[-1]
tot = 0
for s in seq[a:b]:
    print(f"{s[1]:8.2f} us  {s[2]:>14}  {s[0]}")
    tot += s[1]
print('total', round(tot, 1), 'us over', b - a, 'launches')
