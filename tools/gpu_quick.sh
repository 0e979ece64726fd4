#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -6
echo "=== bench"; timeout 600 python bench.py --steps 6 > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("csmri x3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "upd us/it", round(d["roofline_update"]["us_per_iteration"],1), "parity", d.get("parity",{}).get("rel_max_err_vs_oracle"), "floor", d.get("parity",{}).get("fp32_floor"))
print("      fp16", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["fp16"].items()})
for t,r in d["tasks"].items():
    print(t, "x3 value", round(r["value"]), "frac", round(r["roofline"]["frac"],3), "upd us/it", round(r["roofline_update"]["us_per_iteration"],1), "parity", r.get("parity",{}).get("rel_max_err_vs_oracle"), "floor", r.get("parity",{}).get("fp32_floor"), "| fp16", round(r["fp16"]["value"]), round(r["fp16"]["frac"],3), r["fp16"].get("parity"))
PY
