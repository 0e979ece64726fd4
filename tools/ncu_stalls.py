#!/usr/bin/env python
"""Issue / stall summary per kernel from an `ncu --set full` report (reads `ncu -i <rep> --page raw --csv`).

    python tools/ncu_stalls.py /tmp/upd_pr.ncu-rep > gpurun_out/ncu_stalls_pr.txt
Columns: duration, issue-active %, achieved warps/SM, executed warp instructions, L2 %, and the top stall reasons
(warps stalled per issue slot).  Complements tools/ncu_table.py (bytes and pipes)."""
import csv, io, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr, data = rows[hi], rows[hi + 2:]
def f(r, name):
    for i, h in enumerate(hdr):
        if h == name and r[i] not in ("", "no data", "n/a"):
            try: return float(r[i].replace(",", ""))
            except ValueError: pass
    return float("nan")
kn = hdr.index("Kernel Name")
stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in data:
    name = r[kn].replace("void ", "").replace("tfpnp::", "").replace("<unnamed>::", "").split("(")[0]
    st = sorted(((f(r, h), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for h in stalls), reverse=True)
    top = ", ".join(f"{n} {v:.2f}" for v, n in st[:5] if v == v)
    print(f"{name[:40]:40s} {f(r,'gpu__time_duration.sum'):8.1f} us  issue {f(r,'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):5.1f}%  "
          f"warps/SM {f(r,'sm__warps_active.avg.pct_of_peak_sustained_active')*0.64:5.1f}  inst {f(r,'smsp__inst_executed.sum')/1e6:6.1f}M  "
          f"lts {f(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}%  | {top}")
