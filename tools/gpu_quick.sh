#!/bin/bash
# Scratch script for one short gpurun call while iterating on a kernel (edit freely; tools/gpu_final.sh is the round-end check,
# tools/gpu_profiles.sh + tools/profiles_collect.sh regenerate profiles/).  Every command carries its own timeout: a hung kernel
# must not hold the box until gpurun's limit.
#   /usr/local/graft/bin/gpurun --timeout 900 -- 'bash tools/gpu_quick.sh'
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "csmri" 2>&1 | tail -3
for p in fp16x3 fp16; do
  timeout 300 python bench.py --steps 6 --tasks csmri --no-cpu-baseline --precision $p > gpurun_out/bench_q_$p.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_q_$p.json"))
print("$p value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3), "update us/iter", round(d["roofline_update"]["us_per_iteration"], 1))
PY
done
