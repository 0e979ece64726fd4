#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -12
