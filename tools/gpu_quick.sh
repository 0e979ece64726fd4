#!/bin/bash
# scratch: quick check after a change (edit freely); the end-of-block run is tools/gpu_round.sh
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
echo "=== pytest -m gpu"; timeout 600 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
echo "=== smoke"; timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench"; timeout 300 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; cut -c1-1800 gpurun_out/bench.json
