#!/bin/bash
# scratch: quick check after a change (edit freely); the end-of-block run is tools/gpu_round.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest fused"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -k "fused_update or csmri_golden" 2>&1 | tail -4
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|upsample|csmri|psnr|pack|gather_params" -s 100 -c 300 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -24
