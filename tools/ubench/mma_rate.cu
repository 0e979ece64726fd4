// Micro-benchmark: tcgen05.mma issue/execute rate for the shapes the conv kernels use.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../tfpnp_b200/csrc mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "sm100.cuh"
using namespace tfpnp::sm100;

__device__ __forceinline__ uint64_t packd(uint32_t lo, uint32_t hi) {
  uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d;
}

// mode 0: one accumulator, fixed descriptors; 1: two accumulators alternating; 2: one accumulator, descriptors
// advance like the conv (tap/kk offsets); 3: two accumulators + advancing descriptors (the conv pattern)
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) mma_rate(long long* out, int iters, int sbo_rows) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, 128, sbo_rows * 128, 0) >> 32);
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, 128) >> 32);
    const uint32_t a0 = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b0 = (smem_u32(smem + 44 * 1024) >> 4) | (1u << 16);   // (A: 324 rows x 128 B = 41.5 KB)
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_tap = (MODE >= 2) ? a0 + (((tap / 3) * 18 + tap % 3) * 128 >> 4) : a0;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint32_t ko = (MODE >= 2) ? kk * 2 : 0;
            const uint64_t bd = packd(b0 + ko, b_hi);
            umma_f16(tm, packd(a_tap + ko, a_hi), bd, idesc, 1);
            if (MODE == 1 || MODE == 3) umma_f16(tm + N, packd(a_tap + (8 * 128 >> 4) + ko, a_hi), bd, idesc, 1);
            else umma_f16(tm, packd(a_tap + ko, a_hi), bd, idesc, 1);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int N, int MODE>
void run(const char* name, int sbo_rows) {
  long long* d; cudaMalloc(&d, 16);
  const int iters = 50, smem = 100 * 1024;
  cudaFuncSetAttribute(mma_rate<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mma_rate<N, MODE><<<1, 128, smem>>>(d, iters, sbo_rows);
  mma_rate<N, MODE><<<1, 128, smem>>>(d, iters, sbo_rows);
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const double n = iters * 72.0;
  printf("%-34s N=%3d sbo_rows=%2d : issue %.1f clk/MMA, complete %.1f clk/MMA  %s\n", name, N, sbo_rows, h[0] / n, h[1] / n,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  run<64, 0>("1 accum, fixed desc", 8);
  run<64, 1>("2 accum, fixed desc", 8);
  run<64, 2>("1 accum, conv desc pattern", 18);
  run<64, 3>("2 accum, conv desc pattern", 18);
  run<64, 3>("2 accum, conv desc, dense SBO", 8);
  run<32, 3>("2 accum, conv desc pattern", 18);
  run<128, 3>("2 accum, conv desc pattern", 18);
  run<128, 0>("1 accum, fixed desc", 8);
  run<256, 0>("1 accum, fixed desc", 8);
  run<32, 0>("1 accum, fixed desc", 8);
  run<96, 0>("1 accum, fixed desc", 8);
  run<96, 3>("2 accum, conv desc pattern", 18);
  run<192, 0>("1 accum, fixed desc", 8);
  run<192, 3>("2 accum, conv desc pattern", 18);
  run<256, 3>("2 accum, conv desc pattern", 18);
  return 0;
}
