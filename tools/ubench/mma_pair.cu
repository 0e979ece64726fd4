// Micro-benchmark 3 (bring-up for DESIGN 9.1): tcgen05.mma.cta_group::2 on a CTA pair -- numerics + rate.
//   D[256, N] = A[256, 64] * B[N, 64]^T   (fp16 operands, fp32 accumulate), M = 256 split over the pair (128 rows per CTA),
//   B split by rows: CTA r holds B rows [r*N/2, (r+1)*N/2) at the same shared-memory offset.
// Operands are written into shared memory with generic stores in the 128-byte-swizzled K-major layout TMA would produce
// (16-byte chunk index ^= row & 7), the leader (rank 0) issues the MMAs, the completion is multicast to both CTAs, each CTA
// reads its 128 accumulator rows from its own TMEM.  The host checks D against a CPU product and reports clk/MMA.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../tfpnp_b200/csrc mma_pair.cu -o mma_pair
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include "sm100.cuh"
using namespace tfpnp::sm100;

__device__ __forceinline__ uint64_t packd(uint32_t lo, uint32_t hi) {
  uint64_t d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi)); return d;
}
__device__ __forceinline__ void umma2_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

template <int N>
__global__ void __launch_bounds__(128, 1) mma_pair(const __half* __restrict__ A, const __half* __restrict__ B,
                                                   float* __restrict__ D, long long* __restrict__ clk, int iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // [128 rows][128 B]
  uint8_t* sB = smem + 128 * 128;        // [N/2 rows][128 B]
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  // operands -> swizzled K-major tiles (K = 64 halfs = 128 B per row)
  for (int i = threadIdx.x; i < 128 * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(A + ((size_t)(rank * 128 + row) * 64 + ch * 8));
    *reinterpret_cast<uint4*>(sA + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
  }
  for (int i = threadIdx.x; i < (N / 2) * 8; i += 128) {
    const int row = i >> 3, ch = i & 7;
    const uint4 v = *reinterpret_cast<const uint4*>(B + ((size_t)(rank * (N / 2) + row) * 64 + ch * 8));
    *reinterpret_cast<uint4*>(sB + row * 128 + ((ch ^ (row & 7)) << 4)) = v;
  }
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(256) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                     // both CTAs: operands written, barriers initialised, TMEM allocated
  tc_fence_after();
  const uint32_t tm = slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {           // the leader issues for the pair
    const uint32_t idesc = make_idesc_f16(256, N);
    const uint32_t hi = (uint32_t)(make_smem_desc(0, 128) >> 32);
    const uint32_t a0 = (smem_u32(sA) >> 4) | (1u << 16);
    const uint32_t b0 = (smem_u32(sB) >> 4) | (1u << 16);
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma2_f16(tm, packd(a0 + kk * 2, hi), packd(b0 + kk * 2, hi), idesc, (it | kk) != 0);
      }
      __syncwarp();
    }
    if (elect_one()) umma2_commit_mc(&bar, 3);
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  t1 = clock64();
  tc_fence_after();
  if (warp == 1 && rank == 0 && lane == 0) { clk[0] = t1 - t0; }
  // each CTA reads its own 128 rows
  for (int c0 = 0; c0 < N; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32(tm + ((uint32_t)(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    const int row = rank * 128 + warp * 32 + lane;
    for (int j = 0; j < 32; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256) : "memory");
}

template <int N>
int run() {
  std::vector<__half> hA(256 * 64), hB(N * 64);
  std::vector<float> fA(256 * 64), fB(N * 64);
  unsigned s = 12345u;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return ((int)((s >> 9) & 1023) - 512) / 512.0f; };
  for (size_t i = 0; i < hA.size(); ++i) { hA[i] = __float2half(rnd()); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { hB[i] = __float2half(rnd()); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD; long long* dC;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 256 * N * 4); cudaMalloc(&dC, 16);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, 256 * N * 4);
  const int smem = 128 * 128 + (N / 2) * 128 + 2048;
  cudaFuncSetAttribute(mma_pair<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  // 1 iteration: numerics
  cudaError_t e = cudaLaunchKernelEx(&cfg, mma_pair<N>, (const __half*)dA, (const __half*)dB, dD, dC, 1);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: launch/sync failed: %s\n", N, cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(256 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0; int bad_row = -1, bad_col = -1;
  for (int m = 0; m < 256; ++m)
    for (int n = 0; n < N; ++n) {
      double acc = 0;
      for (int k = 0; k < 64; ++k) acc += (double)fA[m * 64 + k] * fB[n * 64 + k];
      const double err = fabs(acc - hD[(size_t)m * N + n]);
      if (err > maxerr) { maxerr = err; bad_row = m; bad_col = n; }
      if (fabs(acc) > maxref) maxref = fabs(acc);
    }
  printf("N=%3d numerics: max |err| %.3e (max |ref| %.2f) at (%d,%d)  %s\n", N, maxerr, maxref, bad_row, bad_col,
         maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  // rate
  e = cudaLaunchKernelEx(&cfg, mma_pair<N>, (const __half*)dA, (const __half*)dB, dD, dC, 2000);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  long long c = 0; cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost);
  printf("N=%3d rate: %.1f clk per M=256 MMA (cta_group::1 M=128 needs max(N/2, 32+N/4) = %d)  %s\n", N, c / (2000.0 * 4),
         (N / 2 > 32 + N / 4) ? N / 2 : 32 + N / 4, e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return 0;
}

int main() {
  run<128>();
  run<64>();
  run<256>();
  return 0;
}
