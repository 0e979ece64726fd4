#!/bin/bash
# scratch: quick check after a change (edit freely); the end-of-block run is tools/gpu_round.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== bench"; timeout 600 python bench.py 2>&1 | tail -1 | cut -c1-2600
