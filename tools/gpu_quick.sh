#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest measure"; timeout 600 python -m pytest tests/test_measure.py -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -15
