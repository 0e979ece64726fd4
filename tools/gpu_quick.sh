#!/bin/bash
mkdir -p gpurun_out
for p in fp16x3 fp16; do
timeout 600 python bench.py --steps 6 --tasks csmri --no-cpu-baseline --precision $p > gpurun_out/bench_q.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("$p value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e"]["ms_per_step"], d["ms_per_step"], "traffic", d["roofline"]["traffic"])
PY
done
