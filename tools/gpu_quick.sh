#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -m gpu -q --timeout 600 -p no:cacheprovider -x -k "denoiser or csmri or conv or cfg2" 2>&1 | tail -3
for p in fp16 fp16x3; do timeout 200 python tools/layer_profile.py --precision $p | tee gpurun_out/layer_profile_$p.txt | grep -E "l12|l13|l14|per call"; done
