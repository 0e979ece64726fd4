#!/bin/bash
# One GPU call at the end of a work block: full -m gpu suite, smoke, bench (own arm + reference arm), the ncu launch
# list of the bench command and `--set full` captures of one denoiser call's kernels.  Numbers printed by runs under
# ncu are never bench values.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== reverse mode (SURVEY 8f N4; written after round 1's GPU budget was spent, gated until it has passed here once)"
TFPNP_TEST_GRAD=1 timeout 900 python -m pytest tests/test_grad.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_grad.log 2>&1
tail -15 gpurun_out/pytest_grad.log
echo "=== reverse mode, layer by layer against the CPU emulation (first deviating region names the kernel)"
for m in 0 2; do TFPNP_GRAD_TC=$m timeout 300 python tools/grad_layer_check.py > gpurun_out/grad_layers_$m.log 2>&1; grep -E "mode|deviates|gx|gsigma" gpurun_out/grad_layers_$m.log | head -12; done
echo "=== reverse-mode timing at the north-star shape (CUDA-core vs tensor-core VJP convolutions)"
timeout 900 python tools/grad_bench.py > gpurun_out/grad_bench.log 2>&1; tail -5 gpurun_out/grad_bench.log
echo "=== reverse mode of all four tasks at their BASELINE per-GPU shapes (split-fp16 tensor-core VJP)"
timeout 900 python tools/grad_tasks.py > gpurun_out/grad_tasks.log 2>&1; tail -6 gpurun_out/grad_tasks.log
echo "=== experiment: row-independent transform warps (TFPNP_XFORM2=1): bit-identity test, then the bench with it"
TFPNP_TEST_XFORM2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k xform2 -p no:cacheprovider 2>&1 | tail -2
TFPNP_TEST_XFORM2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k xform2 -p no:cacheprovider 2>&1 | grep "xform2=2"
for v in 1 2; do TFPNP_XFORM2=$v timeout 600 python bench.py > gpurun_out/bench_xform2_$v.json 2> gpurun_out/bench_xform2_$v.err; cut -c1-220 gpurun_out/bench_xform2_$v.json; grep -o '"parity": {[^}]*}' gpurun_out/bench_xform2_$v.json; done
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-2500 gpurun_out/bench.json
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
cut -c1-900 gpurun_out/bench_reference.json
if [ "$1" != "noncu" ]; then
echo "=== ncu launch list"
# (the bench builds its inputs on the GPU with many small torch kernels first: list this repo's kernels only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv|upsample|csmri|psnr|pack|gather_params" -s 100 -c 300 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -24
echo "=== ncu full: every kernel of one inner iteration"
timeout 900 ncu --set full --clock-control none -k regex:"conv|upsample|csmri_rows|csmri_cols" -s 64 -c 32 \
  -o gpurun_out/iter_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
echo "=== ncu full + source: the three heaviest conv variants"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k regex:"conv3x3_tc2<.int.32, .int.32, .bool.1, .bool.1|conv3x3_tc2<.int.128, .int.64, .bool.0, .bool.0, .bool.0|conv3x3_tc2<.int.32, .int.32, .bool.1, .bool.0" \
  -s 12 -c 3 -o gpurun_out/conv_src -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_src.log 2>&1
tail -1 gpurun_out/ncu_src.log
timeout 300 ncu --set full --clock-control none -k regex:"spi_" -s 1 -c 2 -o gpurun_out/upd_spi -f python tools/run_tasks.py spi > gpurun_out/ncu_upd_spi.log 2>&1
tail -1 gpurun_out/ncu_upd_spi.log
ls -la gpurun_out/*.ncu-rep
fi
