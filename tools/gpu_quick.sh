#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_env.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 -p no:cacheprovider -k "denoiser or csmri" 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-1400 | tee gpurun_out/bench_quick.json
echo "=== ncu fused"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv3x3_tc2<64, 64, 0, 1|conv3x3_tc2<32, 32, 1, 1" -s 8 -c 2 \
  -o gpurun_out/conv_fuse -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_fuse.log 2>&1
tail -2 gpurun_out/ncu_fuse.log
