#!/bin/bash
# Exploratory GPU run: each group in its own process so a sticky CUDA error in one
# group (e.g. a trapped tensor-core kernel) cannot poison the next.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name" ; timeout 900 python -m pytest "$@" -q --timeout 300 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/$name.log; tail -5 gpurun_out/$name.log; }
run conv     tests/test_gpu_conv.py -m gpu
run simt     tests/test_gpu_parity.py -m gpu -k "fp32_simt or mask_selection or radon or psnr or spi_prox or library"
run x3       tests/test_gpu_parity.py -m gpu -k "fp16x3 or tc_matches or call_semantics or pr_vs_oracle"
run fp16     tests/test_gpu_parity.py -m gpu -k "fp16 and not fp16x3"
