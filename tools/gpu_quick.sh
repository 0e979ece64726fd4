#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest conv"; timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q -x --timeout 120 -p no:cacheprovider 2>&1 | tail -6
echo "=== pytest denoiser"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "denoiser or csmri" 2>&1 | tail -6
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-300
echo "=== bench rows off"; TFPNP_CONV_ROWS=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-300
echo "=== knockout"; timeout 300 python tools/conv_knockout.py 2>&1 | head -1 | cut -c1-400
