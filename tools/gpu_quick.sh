#!/bin/bash
# scratch: quick check after a change (edit freely)
mkdir -p gpurun_out
echo "=== pytest denoiser"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "denoiser or conv" 2>&1 | tail -8
echo "=== bench x3"; timeout 300 python bench.py --precision fp16x3 --no-cpu-baseline > gpurun_out/bench_x3.json 2>gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_x3.json; tail -3 gpurun_out/bench.err
echo "=== bench fp16"; timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_fp16.json 2>gpurun_out/bench.err; cut -c1-300 gpurun_out/bench_fp16.json; tail -3 gpurun_out/bench.err
