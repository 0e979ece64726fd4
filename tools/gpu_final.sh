#!/bin/bash
# Round-end check on one GPU: the whole -m gpu suite, smoke(), and both bench arms (bounded by their own timeouts).
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 540 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -2 gpurun_out/bench.err
timeout 240 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; echo "reference rc=$?"
wc -l gpurun_out/bench.json gpurun_out/bench_reference.json
