#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/fuse_probe.py | grep -E "B 12|B 5|repeat|oracle"
echo "=== pytest"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider -x 2>&1 | tail -5
timeout 200 python tools/layer_profile.py --precision fp16x3 | tee gpurun_out/layer_profile_fp16x3.txt | grep -E "l24|l21|l18|upsample|per call"
echo "=== bench"; timeout 600 python bench.py --steps 4 --tasks csmri > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("csmri x3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), "parity", d.get("parity",{}).get("rel_max_err_vs_oracle"), "floor", d.get("parity",{}).get("fp32_floor"), d["clocks"])
PY
