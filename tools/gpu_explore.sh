#!/bin/bash
mkdir -p gpurun_out
echo "=== mma_chip"; timeout 300 tools/ubench/mma_chip 2>&1 | tee gpurun_out/mma_chip.txt
echo "=== knockout (graph-timed)"; timeout 900 python tools/conv_knockout.py 2>&1 | tee gpurun_out/knockout.txt
