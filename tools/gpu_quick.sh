#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "csmri or denoiser" 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-1500
echo "=== bench unfused csmri"; TFPNP_CSMRI_FUSED=0 timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-300
