#!/bin/bash
# scratch: quick check after a change (edit freely); the end-of-block run is tools/gpu_round.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest (conv + solver parity)"; timeout 900 python -m pytest tests/test_gpu_conv.py tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
echo "=== bench pair"; TFPNP_CONV_PAIR=1 timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 | cut -c1-300
