#!/bin/bash
# quick check after a kernel change: conv + denoiser parity tests, bench (no CPU baseline), knock-out timings
mkdir -p gpurun_out
echo "=== pytest conv+denoiser"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -p no:cacheprovider 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-1400 | tee gpurun_out/bench_quick.json
echo "=== knockout"; timeout 600 python tools/conv_knockout.py 2>&1 | head -3 | tee gpurun_out/knockout.txt
