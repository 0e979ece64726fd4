#!/bin/bash
# One GPU call: full -m gpu suite, smoke, bench (own arm), launch list + one `--set full` capture of the
# top conv kernels.  Numbers printed by the runs under ncu are never bench values.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== pytest -m gpu"
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench"
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-1500 gpurun_out/bench.json
if [ "$1" != "noncu" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 340 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
python tools/launch_summary.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -24
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_tc2 -s 30 -c 24 \
  -o gpurun_out/conv_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
fi
