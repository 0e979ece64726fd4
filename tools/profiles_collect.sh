#!/bin/bash
# Turns the captures tools/gpu_profiles.sh left in gpurun_out/ into the tracked summaries under profiles/ (runs here, no GPU).
R=${1:-r02}
SHA=$(cat gpurun_out/src_sha16.txt)
cp gpurun_out/bench.json profiles/${R}_bench.json
cp gpurun_out/bench_reference.json profiles/${R}_bench_reference.json
for p in fp16x3 fp16; do
  cp gpurun_out/layer_profile_$p.txt profiles/${R}_layer_profile_$p.txt
  cp gpurun_out/launches_$p.csv profiles/${R}_launches_$p.csv
  cp gpurun_out/launches_${p}_summary.txt profiles/${R}_launches_${p}_summary.txt
  cp gpurun_out/ncu_iter_full_$p.txt profiles/${R}_ncu_iter_full_$p.txt
  cp gpurun_out/traffic_$p.json /tmp/traffic_$p.json
  python - <<PY
import json
d = json.load(open("/tmp/traffic_$p.json"))
d["src_sha16"] = "$SHA"
d["note"] = "one inner iteration (a window of exactly one period of the launch sequence): one denoiser call + one update; cold-cache replays"
json.dump(d, open("profiles/${R}_traffic_csmri_$p.json", "w"))
PY
done
for t in csmri pr ct spi; do echo "== $t"; cat gpurun_out/ncu_upd_$t.txt; done > profiles/${R}_update_kernels.txt
cp gpurun_out/gpu.txt profiles/${R}_gpu.txt
ls -la profiles/${R}_*
