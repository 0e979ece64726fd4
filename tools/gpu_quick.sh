#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests/test_grad.py tests/test_csmri_variants.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/grad_bench.py 2>&1 | tail -8
