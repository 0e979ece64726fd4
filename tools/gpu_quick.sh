#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "denoiser or csmri" 2>&1 | tail -4
timeout 200 python tools/layer_profile.py --precision fp16x3 | tee gpurun_out/layer_profile_fp16x3.txt | grep -E "pair|per call"
echo "=== bench"; timeout 600 python bench.py --steps 4 --tasks csmri --no-cpu-baseline > gpurun_out/bench.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench.json"))
print("csmri x3 value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"],3), d["clocks"])
PY
