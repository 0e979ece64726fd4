#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "denoiser or csmri" 2>&1 | tail -3
for v in 0 1; do
echo "=== FIRST_LATE=$v"; TFPNP_FIRST_LATE=$v timeout 600 python bench.py --steps 6 --tasks csmri --no-cpu-baseline > gpurun_out/bench_$v.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$v.json"))
print("csmri x3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"])
print("   fp16", round(d["fp16"]["value"]), round(d["fp16"]["frac"],3))
PY
done
