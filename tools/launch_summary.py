#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by (kernel, grid)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]
ki, vi, gi = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Grid Size')
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    name = r[ki].replace('unnamed>::', '').replace('void ', '')[:48]
    a = agg.setdefault((name, r[gi]), [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
tot = sum(a[1] for a in agg.values())
print(f"{'total us':>10} {'n':>4} {'avg us':>8} {'share':>6}  kernel, grid")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{a[1]/1e3:10.1f} {a[0]:4d} {a[1]/a[0]/1e3:8.1f} {100*a[1]/tot:5.1f}%  {k[0]} {k[1]}")
print(f"total {tot/1e6:.3f} ms over {sum(a[0] for a in agg.values())} launches")
