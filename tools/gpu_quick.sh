#!/bin/bash
# scratch: quick check after a change (edit freely); the end-of-block run is tools/gpu_round.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
for sa in 3 2; do echo "=== TFPNP_SMALL_SA=$sa"; TFPNP_SMALL_SA=$sa timeout 300 python tools/conv_knockout.py 2>&1 | head -1 | cut -c1-330; done
