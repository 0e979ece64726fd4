#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 6 --tasks csmri --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("csmri x3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], "| fp16", round(d["fp16"]["value"]), round(d["fp16"]["frac"],3))
PY
