#!/usr/bin/env python
"""Top warp-stall source lines from `ncu -i rep --page source --csv --print-source cuda,sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
pat = sys.argv[2] if len(sys.argv) > 2 else ''
secs = []; cur = None
for r in rows:
    if r and r[0] == 'File Path': cur = {'file': r[1], 'rows': []}; secs.append(cur)
    elif r and r[0] == 'Function Name': cur['func'] = r[1]
    elif r and r[0] == 'Line No': cur['hdr'] = r
    elif cur is not None and r: cur['rows'].append(r)
def num(x):
    try: return int(x)
    except Exception: return 0
seen = set()
for s in secs:
    key = (s['file'], s.get('func'))
    if key in seen or pat not in s.get('func', ''): continue
    seen.add(key)
    h = s['hdr']; si = h.index('# Samples')
    tot = sum(num(r[si]) for r in s['rows'])
    print(s['file'].split('/')[-1], s['func'][:70], 'samples', tot)
    for r in sorted(s['rows'], key=lambda r: -num(r[si]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
        if num(r[si]): print(f"   {r[0]:>5} {num(r[si]):6d}  {r[1][:120]}")
