#!/bin/bash
# round 2, call A: run what round 1 wrote blind (reverse mode, XFORM2), and baseline the fp16x3 mode on the shipped kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
cat gpurun_out/gpu.txt
echo "=== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -4 gpurun_out/pytest_gpu.log
echo "=== reverse mode"
TFPNP_TEST_GRAD=1 timeout 900 python -m pytest tests/test_grad.py -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_grad.log 2>&1
tail -30 gpurun_out/pytest_grad.log
echo "=== reverse mode, layer by layer"
for m in 0 2; do TFPNP_GRAD_TC=$m timeout 300 python tools/grad_layer_check.py > gpurun_out/grad_layers_$m.log 2>&1; grep -E "mode|deviates|gx|gsigma" gpurun_out/grad_layers_$m.log | head -12; done
echo "=== bench fp16 (default)"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_fp16.json; echo
echo "=== bench fp16x3 on the shipped kernels"
timeout 300 python bench.py --precision fp16x3 --no-cpu-baseline > gpurun_out/bench_x3.json 2>> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_x3.json; echo
echo "=== XFORM2"
TFPNP_TEST_XFORM2=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k xform2 -p no:cacheprovider 2>&1 | tail -3
for v in 1 2; do TFPNP_XFORM2=$v timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_xform2_$v.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_xform2_$v.json; echo; done
echo "=== FUSE_UP_MIN=64"
TFPNP_FUSE_UP_MIN=64 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_fuseup64.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_fuseup64.json; echo
echo "=== reverse-mode timing"
timeout 600 python tools/grad_bench.py > gpurun_out/grad_bench.log 2>&1; tail -5 gpurun_out/grad_bench.log
echo "=== sanitize"
timeout 600 bash tools/sanitize.sh > gpurun_out/sanitize.log 2>&1; tail -8 gpurun_out/sanitize.log
