#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_csmri_variants.py -m gpu -x -q -k "csmri" 2>&1 | tail -3
for pdl in 7 0; do
for p in fp16x3 fp16; do
TFPNP_CSMRI_PDLMASK=$pdl timeout 600 python bench.py --steps 6 --tasks csmri --no-cpu-baseline --precision $p > gpurun_out/bench_q.json 2>gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json"))
print("pdlmask $pdl $p value", round(d["value"]), "update us/iter", round(d.get("roofline_update",{}).get("us_per_iteration"),1), "e2e", round(d["e2e"]["value"]))
PY
done
done
