#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do
echo "=== PAIR_SINGLE=$v"; TFPNP_PAIR_SINGLE=$v timeout 600 python bench.py --steps 4 --tasks csmri --no-cpu-baseline > gpurun_out/bench_s$v.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_s$v.json"))
print("csmri x3 value", round(d["value"]), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"]), "| fp16", round(d["fp16"]["value"]), round(d["fp16"]["e2e"]))
PY
done
TFPNP_PAIR_SINGLE=1 timeout 200 python tools/layer_profile.py --precision fp16x3 | grep -E "pair|per call"
