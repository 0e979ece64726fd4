#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -p no:cacheprovider -k "csmri or pr_" 2>&1 | tail -4
for t in csmri pr; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"${t}_cols|${t}_rows" -s 6 -c 3 python tools/run_tasks.py $t 2>&1 | grep -E "gpu__time|_cols|_rows" | grep -v PROF | cut -c1-60
done
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | cut -c1-300
