#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest ct"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_env.py -m gpu -q --timeout 600 -p no:cacheprovider -k "radon or ct" 2>&1 | tail -5
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"radon|ct_|transpose" -s 8 -c 6 python tools/run_tasks.py ct 2>&1 | grep -E "radon_fwd|ct_bwd|transpose|gpu__time" | head -14
