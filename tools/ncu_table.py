#!/usr/bin/env python
"""Per-kernel table from an `ncu --set full` report: duration, tensor-pipe %, DRAM %, DRAM bytes, L2 bytes, registers.

    python tools/ncu_table.py gpurun_out/iter_full.ncu-rep [--period N] [--json out.json] > profiles/rNN_ncu_iter_full.txt

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU).  `--period N` keeps the first N launches
(= one inner iteration).  The JSON side file carries the DRAM bytes per iteration split into denoiser / update kernels
(what bench.py prints as `roofline.traffic` -- only when regenerated from the shipped binary in the same round).
The dram% column is dram__throughput.avg.pct_of_peak_sustained_elapsed when the report has it, else (read + write bytes) /
duration over the measured copy bandwidth in MEASURED_PEAKS.json.
Numbers under ncu are cold-cache, serialised replays: compare shares and percentages, not absolutes."""
import argparse
import csv
import io
import json
import subprocess

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("--period", type=int, default=0)
ap.add_argument("--period-from", default=None, help="regex: keep the launches from the first matching kernel up to (not including) its next occurrence")
ap.add_argument("--json", default=None)
ap.add_argument("--update-regex", default="csmri|pr_|pr256|ct_|radon|pad_transpose|spi_")
a = ap.parse_args()

raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
if a.period_from:
    import re as _re
    kn = hdr.index("Kernel Name")
    hits = [i for i, r in enumerate(data) if _re.search(a.period_from, r[kn])]
    if len(hits) >= 2:
        data = data[hits[0]:hits[1]]
elif a.period:
    data = data[:a.period]


def cols(name):
    exact = [i for i, h in enumerate(hdr) if h == name]
    return exact + [i for i, h in enumerate(hdr) if h.endswith("." + name)]


def col(name):
    c = cols(name)
    return c[0] if c else None


def val(r, name, scale_units=None):
    for i in cols(name):          # (the "TriageCompute" copies of a metric are empty unless that section was collected)
        if i < len(r) and r[i] not in ("", "no data", "n/a"):
            try:
                v = float(r[i].replace(",", ""))
            except ValueError:
                continue
            if scale_units:
                v *= scale_units.get(units[i], 1)
            return v
    return float("nan")


try:
    import os as _os
    HBM_GBS = json.load(open(_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    HBM_GBS = 6549.8
BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
US = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}
import re
upd_re = re.compile(a.update_regex)
print(f"{'kernel':46s} {'grid':>10s} {'us':>7s} {'tensor%':>8s} {'dram%':>6s} {'dramR MB':>9s} {'dramW MB':>9s} {'L2 MB':>8s} {'lts%':>5s} {'regs':>4s} {'smem KB':>7s}")
tot = den = upd = 0.0
den_us = upd_us = 0.0
for r in data:
    name = r[col("Kernel Name")].replace("void ", "").replace("tfpnp::", "").replace("<unnamed>::", "").split("(")[0]
    t = val(r, "gpu__time_duration.sum", US)
    tens = val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
    if tens != tens:
        tens = val(r, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed")
    dram = val(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed")
    dr, dw = val(r, "dram__bytes_read.sum", BYTES), val(r, "dram__bytes_write.sum", BYTES)
    if dram != dram and t > 0:          # (this ncu's --set full has no dram__throughput pct: bytes / time over the measured copy bandwidth)
        dram = 100.0 * (dr + dw) / (t * 1e-6) / (HBM_GBS * 1e9)
    l2 = val(r, "lts__t_sectors.sum") * 32
    lts = val(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed")
    regs = val(r, "launch__registers_per_thread")
    smem = val(r, "launch__shared_mem_per_block_dynamic", BYTES) / 1e3
    if smem < 1:                       # some ncu versions report the column in Kbyte without a unit row entry
        smem = val(r, "launch__shared_mem_per_block_dynamic")
    grid = r[col("Grid Size")].replace(" ", "")
    tot += t
    if upd_re.search(name):
        upd += dr + dw; upd_us += t
    else:
        den += dr + dw; den_us += t
    print(f"{name[:46]:46s} {grid:>10s} {t:7.1f} {tens:8.1f} {dram:6.1f} {dr/1e6:9.2f} {dw/1e6:9.2f} {l2/1e6:8.1f} {lts:5.1f} {regs:4.0f} {smem:7.1f}")
print(f"{len(data)} launches, {tot:.1f} us serialised (cold-cache replays): denoiser kernels {den_us:.1f} us / {den/1e6:.1f} MB DRAM, "
      f"update kernels {upd_us:.1f} us / {upd/1e6:.1f} MB DRAM")
if a.json:
    json.dump({"denoiser_dram_bytes_per_iter": den, "update_dram_bytes_per_iter": upd, "launches_per_iter": len(data),
               "source": f"ncu --set full --clock-control none ({a.report}), tools/ncu_table.py"}, open(a.json, "w"))
