// Micro-benchmark 2: tcgen05.mma pace with the WHOLE chip busy (grid = SMs x CTAs/SM), for 128-byte and 64-byte
// operand rows, zero vs random operand data, with the real SM clock (clock64 / globaltimer) reported.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../tfpnp_b200/csrc mma_chip.cu -o mma_chip
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "sm100.cuh"
using namespace tfpnp::sm100;

__device__ __forceinline__ uint64_t packd(uint32_t lo, uint32_t hi) {
  uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d;
}
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}

// Each CTA: 2 accumulators, conv descriptor pattern (9 taps x KSTEPS), `iters` repetitions.
template <int N, int ROWB, int COMMIT>
__global__ void __launch_bounds__(128) mma_chip(long long* out, int iters, int rnd, int tmem_cols) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t sink[4];
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int KSTEPS = ROWB / 32;
  constexpr int A_BYTES = ((324 * ROWB + 1023) / 1024) * 1024;
  constexpr int TOTAL = A_BYTES + N * ROWB;
  for (int i = threadIdx.x; i < TOTAL / 2; i += 128) {
    unsigned h = (i * 2654435761u + blockIdx.x * 40503u) >> 7;
    reinterpret_cast<__half*>(smem)[i] = rnd ? __float2half(((int)(h & 1023) - 512) * (1.f / 512.f)) : __float2half(0.f);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 4; ++i) mbar_init(&sink[i], 1 << 20); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, tmem_cols);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = slot;
  if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a_hi = (uint32_t)(make_smem_desc_ex(0, ROWB, 18 * ROWB, 0) >> 32);
    const uint32_t b_hi = (uint32_t)(make_smem_desc(0, ROWB) >> 32);
    const uint32_t a0 = (smem_u32(smem) >> 4) | (1u << 16);
    const uint32_t b0 = (smem_u32(smem + A_BYTES) >> 4) | (1u << 16);
    unsigned long long g0 = gtimer();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t a_tap = a0 + (((tap / 3) * 18 + tap % 3) * ROWB >> 4);
#pragma unroll
          for (int kk = 0; kk < KSTEPS; ++kk) {
            const uint64_t bd = packd(b0 + kk * 2, b_hi);
            umma_f16(tm, packd(a_tap + kk * 2, a_hi), bd, idesc, 1);
            umma_f16(tm + N, packd(a_tap + (8 * ROWB >> 4) + kk * 2, a_hi), bd, idesc, 1);
          }
          if (COMMIT == 3 && tap % 3 == 2) umma_commit(&sink[tap / 3]);     // a commit per kernel row (3 taps)
        }
        if (COMMIT >= 1) umma_commit(&sink[0]);                             // a commit per 9-tap block
        if (COMMIT == 2) umma_commit(&sink[1]);
      }
      if (COMMIT == 4) { __syncwarp(); mbar_try_wait(&bar, 1); tc_fence_after(); }   // + an (already satisfied) wait + fence
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    unsigned long long g2 = gtimer();
    if (lane == 0) { out[2 * blockIdx.x] = t2 - t0; out[2 * blockIdx.x + 1] = (long long)(g2 - g0); }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, tmem_cols);
}

template <int N, int ROWB, int COMMIT = 0>
void run(int ctas_per_sm, int rnd) {
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm;
  long long* d; cudaMalloc(&d, 16 * grid);
  const int iters = 400;
  const int smem = (((324 * ROWB + 1023) / 1024) * 1024) + N * ROWB + 2048;
  const int tmem_cols = 2 * N < 32 ? 32 : (2 * N <= 64 ? 64 : (2 * N <= 128 ? 128 : (2 * N <= 256 ? 256 : 512)));
  cudaFuncSetAttribute(mma_chip<N, ROWB, COMMIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 3; ++rep) mma_chip<N, ROWB, COMMIT><<<grid, 128, smem>>>(d, iters, rnd, tmem_cols);
  cudaError_t e = cudaDeviceSynchronize();
  long long* h = (long long*)malloc(16 * grid);
  cudaMemcpy(h, d, 16 * grid, cudaMemcpyDeviceToHost);
  double clk = 0, ns = 0;
  for (int i = 0; i < grid; ++i) { clk += h[2 * i]; ns += h[2 * i + 1]; }
  clk /= grid; ns /= grid;
  const double n = iters * 9.0 * (ROWB / 32) * 2;
  printf("N=%3d row=%3dB ctas/SM=%d data=%s commit-mode=%d : %.1f clk/MMA per CTA (%.1f per SM), %.1f ns/MMA, SM clock %.0f MHz  %s\n", N, ROWB,
         ctas_per_sm, rnd ? "rand" : "zero", COMMIT, clk / n, clk / n / ctas_per_sm, ns / n, clk / ns * 1e3,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  free(h); cudaFree(d);
}

int main() {
  // commit cost: mode 1 = one tcgen05.commit per 9-tap block, 2 = two, 3 = one per 3 taps + one per block, 4 = 1 + wait + fence
  run<32, 64, 0>(1, 1); run<32, 64, 1>(1, 1); run<32, 64, 2>(1, 1); run<32, 64, 3>(1, 1); run<32, 64, 4>(1, 1);
  run<64, 128, 0>(1, 1); run<64, 128, 1>(1, 1); run<64, 128, 2>(1, 1); run<64, 128, 3>(1, 1); run<64, 128, 4>(1, 1);
  run<128, 128, 0>(1, 1); run<128, 128, 3>(1, 1);
  for (int rnd = 1; rnd < 1; ++rnd) {
    run<32, 64>(1, rnd);
    run<32, 64>(2, rnd);
    run<32, 128>(1, rnd);
    run<64, 64>(1, rnd);
    run<64, 128>(1, rnd);
    run<64, 128>(2, rnd);
    run<128, 128>(1, rnd);
    run<256, 128>(1, rnd);
  }
  return 0;
}
