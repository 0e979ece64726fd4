#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of one inner iteration (tools/gpu_round.sh: iter_full.ncu-rep, exported with
`ncu -i ... --page raw --csv`): per-kernel duration, DRAM and L2 bytes -> a text table and profiles/r01_traffic.json."""
import csv, json, sys
src, out_txt, out_json = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
def g(r, n):
    try:
        return float(r[hdr.index(n)].replace(',', ''))
    except Exception:
        return float('nan')
def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
ur, ut = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('gpu__time_duration.sum')]
lines, den, upd, seen_first = [], 0.0, 0.0, 0
data = rows[2:]
# one inner iteration = `period` consecutive launches (argv[4]; the capture may start anywhere in the cycle)
period = int(sys.argv[4]) if len(sys.argv) > 4 else len(data)
sel = data[:period]
tot_us = 0.0
lines.append(f"{'kernel':44s} {'grid':14s} {'us':>7s} {'dramR MB':>9s} {'dramW MB':>9s} {'L2 MB':>8s} {'regs':>4s}")
for r in sel:
    name = r[hdr.index('Kernel Name')].replace('void ', '').replace('unnamed>::', '').split('(')[0]
    t = to_us(g(r, 'gpu__time_duration.sum'), ut)
    dr, dw = to_bytes(g(r, 'dram__bytes_read.sum'), ur), to_bytes(g(r, 'dram__bytes_write.sum'), units[hdr.index('dram__bytes_write.sum')])
    l2 = g(r, 'lts__t_sectors.sum') * 32
    tot_us += t
    if 'csmri' in name:
        upd += dr + dw
    else:
        den += dr + dw
    lines.append(f"{name[:44]:44s} {r[hdr.index('Grid Size')]:14s} {t:7.1f} {dr/1e6:9.2f} {dw/1e6:9.2f} {l2/1e6:8.1f} {r[hdr.index('launch__registers_per_thread')]:>4s}")
lines.append(f"one inner iteration: {len(sel)} launches, {tot_us:.1f} us (serialised, cold-cache replays); DRAM denoiser {den/1e6:.1f} MB, update {upd/1e6:.1f} MB")
open(out_txt, 'w').write("\n".join(lines) + "\n")
json.dump({"denoiser_dram_bytes_per_iter": den, "update_dram_bytes_per_iter": upd, "launches_per_iter": len(sel),
           "source": "ncu --set full --clock-control none, python bench.py --steps 1 --warmup 3 (tools/gpu_round.sh); " + out_txt},
          open(out_json, 'w'))
print("\n".join(lines[-6:]))
