#!/bin/bash
mkdir -p gpurun_out
for bo in 0 1; do
  echo "=== conv tests TFPNP_DESC_BO=$bo"
  TFPNP_DESC_BO=$bo timeout 600 python -m pytest tests/test_gpu_conv.py -m gpu -q --timeout 120 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/conv_bo$bo.log
  tail -12 gpurun_out/conv_bo$bo.log
done
