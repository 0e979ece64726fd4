#!/bin/bash
# Regenerates everything under profiles/ for the shipped binary (one GPU call; ~6 min).  Numbers printed by runs under ncu are
# never bench values: bench values come from bench.py (CUDA events, max over ranks).
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/gpu_profiles.sh'   then   bash tools/profiles_collect.sh  (here)
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep gpurun_out/*.csv
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== bench (own arm, all tasks) + reference arm"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
echo "=== per-launch in-situ profile (no profiler attached)"
for p in fp16x3 fp16; do timeout 200 python tools/layer_profile.py --precision $p > gpurun_out/layer_profile_$p.txt; tail -1 gpurun_out/layer_profile_$p.txt; done
for p in fp16x3 fp16; do
echo "=== ncu launch list, $p"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|upsample|maxpool|outc|csmri|psnr|pack|gather_params" -s 100 -c 120 --csv \
  --log-file gpurun_out/launches_$p.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tasks csmri --precision $p > gpurun_out/ncu_b_$p.log 2>&1
python tools/launch_summary.py gpurun_out/launches_$p.csv > gpurun_out/launches_${p}_summary.txt; tail -1 gpurun_out/launches_${p}_summary.txt
echo "=== ncu --set full: every kernel of one inner iteration, $p"
# one inner iteration = the launches from one first-layer kernel to the next (denoiser layers + up-samplings + 3 update kernels)
timeout 900 ncu --set full --clock-control none -k regex:"conv|upsample|csmri_rows|csmri_cols" -s 70 -c 72 \
  -o /tmp/iter_full_$p -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --tasks csmri --precision $p > gpurun_out/ncu_full_$p.log 2>&1
tail -1 gpurun_out/ncu_full_$p.log
# (the .ncu-rep files are ~55 MB each: summarised here, only the tables travel back)
python tools/ncu_table.py /tmp/iter_full_$p.ncu-rep --period-from conv_first --json gpurun_out/traffic_$p.json > gpurun_out/ncu_iter_full_$p.txt
tail -1 gpurun_out/ncu_iter_full_$p.txt
done
echo "=== ncu --set full: update kernels of the other tasks"
for t in pr ct spi; do
timeout 300 ncu --set full --clock-control none -k regex:"pr_|pr256|radon|ct_|transpose|spi_" -s 3 -c 6 -o /tmp/upd_$t -f python tools/run_tasks.py $t > gpurun_out/ncu_upd_$t.log 2>&1
tail -1 gpurun_out/ncu_upd_$t.log
python tools/ncu_table.py /tmp/upd_$t.ncu-rep > gpurun_out/ncu_upd_$t.txt
python tools/ncu_stalls.py /tmp/upd_$t.ncu-rep >> gpurun_out/ncu_upd_$t.txt
done
timeout 300 ncu --set full --clock-control none -k regex:"csmri_rows|csmri_cols" -s 3 -c 3 -o /tmp/upd_csmri -f python tools/run_tasks.py csmri > gpurun_out/ncu_upd_csmri.log 2>&1
python tools/ncu_table.py /tmp/upd_csmri.ncu-rep > gpurun_out/ncu_upd_csmri.txt
python tools/ncu_stalls.py /tmp/upd_csmri.ncu-rep >> gpurun_out/ncu_upd_csmri.txt
python tools/src_hash.py > gpurun_out/src_sha16.txt
du -sh gpurun_out
