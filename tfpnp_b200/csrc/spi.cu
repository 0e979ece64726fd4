// Single-photon imaging: prox of the quanta-image-sensor likelihood fused with the
// ADMM dual update (tasks/spi/solver.py:32-44, tfpnp/utils/transforms.py:404-439).
//
// The reference runs ~100 small masked-index launches per iteration (with host syncs
// from boolean indexing); here the whole per-pixel closed form / 10-step bisection
// lives in registers.  Pointwise, HBM-streaming: reads x, u, x0 (12 B/px), writes z, u, d.
// The residual f(y) is evaluated with the reference's operation order and without FMA
// contraction so the bisection takes the same branches as the fp32 CPU code
// (up to the last-ulp difference of expf).
#include "tasks.cuh"

namespace tfpnp {
namespace {

__device__ __forceinline__ float spi_prox(float ztilde, float K1, float Ksq, float mu) {
  const float K0 = __fsub_rn(Ksq, K1);
  float z;
  if (K1 == 0.f) {
    z = __fsub_rn(ztilde, __fdiv_rn(K0, mu));              // transforms.py:415
  } else {
    float bmin = 1e-5f, bmax = 1.1f;
    float bave = __fdiv_rn(__fadd_rn(bmin, bmax), 2.0f);
    bool live = true;
    const float muz = __fmul_rn(mu, ztilde);
#pragma unroll
    for (int i = 0; i < 10; ++i) {                          // transforms.py:427-437
      float e = __fsub_rn(expf(bave), 1.0f);
      float f = __fadd_rn(__fsub_rn(__fsub_rn(__fdiv_rn(K1, e), __fmul_rn(mu, bave)), K0), muz);
      if (live) {
        if (f > 0.f) bmin = bave;
        else if (f < 0.f) bmax = bave;
        else live = false;                                  // f == 0 (or NaN never >,<): frozen
        if (live) bave = __fdiv_rn(__fadd_rn(bmin, bmax), 2.0f);
      }
    }
    z = bave;
  }
  return fminf(fmaxf(z, 0.f), 1.f);
}

__global__ void __launch_bounds__(256)
spi_update_kernel(const float4* __restrict__ x, float4* __restrict__ z, float4* __restrict__ u,
                  float4* __restrict__ d, const float4* __restrict__ x0, const float* __restrict__ K10,
                  const float* __restrict__ mu, int HW4) {
  const int b = blockIdx.y;
  const float K = K10[b];                 // K tensor * 10 (solver.py:32)
  const float Ksq = __fmul_rn(K, K);
  const float m = mu[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW4; i += gridDim.x * blockDim.x) {
    size_t g = (size_t)b * HW4 + i;
    float4 xv = x[g], uv = u[g], cv = x0[g], zv, dv;
    const float* xp = &xv.x; float* up = &uv.x; const float* cp = &cv.x; float* zp = &zv.x; float* dp = &dv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float K1 = __fmul_rn(cp[k], Ksq);                     // K1 = x0 * K^2 (solver.py:33)
      float zz = spi_prox(__fadd_rn(xp[k], up[k]), K1, Ksq, m);
      float un = __fsub_rn(__fadd_rn(up[k], xp[k]), zz);    // u = u + x - z (solver.py:44)
      zp[k] = zz;
      up[k] = un;
      dp[k] = __fsub_rn(zz, un);                            // z - u (solver.py:47)
    }
    z[g] = zv; u[g] = uv; d[g] = dv;
  }
}

}  // namespace

int spi_update(const float* x, float* z, float* u, float* d, const float* x0, const float* K10,
               const float* mu, int B, int HW, cudaStream_t st) {
  TFPNP_CHECK(HW % 4 == 0, "spi: H*W must be a multiple of 4");
  int HW4 = HW / 4;
  int bx = cdiv(HW4, 256);
  spi_update_kernel<<<dim3(bx, B), 256, 0, st>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(z), reinterpret_cast<float4*>(u),
      reinterpret_cast<float4*>(d), reinterpret_cast<const float4*>(x0), K10, mu, HW4);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tfpnp
