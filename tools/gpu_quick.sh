#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 900 python -m pytest tests/test_ircnn.py tests/test_gpu_conv.py -m gpu -q --timeout 600 -p no:cacheprovider 2>&1 | tail -12
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-300
