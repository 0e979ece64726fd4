#!/bin/bash
# Memory-safety check of the reverse mode's index arithmetic WITHOUT a GPU: the layer / iteration sequences, the workspace
# layout and the per-element bodies are shared verbatim between the CUDA kernels and the g++ emulation (grad_elem.cuh), so
# running the emulation under AddressSanitizer + UBSan with exactly-sized buffers checks the same offsets the kernels use.
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/asan
g++ -O1 -g -std=c++17 -shared -fPIC -fsanitize=address,undefined -fno-omit-frame-pointer -o gpurun_out/asan/emu_asan.so tests/grad_elem_host.cpp
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" ASAN_OPTIONS=detect_leaks=0 \
  python tools/asan_emulation.py 2>&1 | grep -v "^$" | tail -20
