#!/bin/bash
# round 2, call B: launch lists (fp16 and fp16x3) + --set full of one inner iteration of the shipped fp16 path
mkdir -p gpurun_out
for prec in fp16 fp16x3; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|upsample|maxpool|outc|csmri|psnr|pack|gather_params" -s 100 -c 120 --csv \
  --log-file gpurun_out/launches_$prec.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --precision $prec > gpurun_out/ncu_b_$prec.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$prec.csv | tee gpurun_out/launches_${prec}_summary.txt | head -30
done
timeout 900 ncu --set full --clock-control none -k regex:"conv|upsample|csmri_rows|csmri_cols" -s 64 -c 32 \
  -o gpurun_out/iter_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
