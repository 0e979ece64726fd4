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

// ---- reverse mode: sequence and element bodies in grad_elem.cuh -----------------------------------------
namespace {

__global__ void spi_slot_copy(const float* __restrict__ state, float* __restrict__ buf, int k, int HW, size_t n, int to_state) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* s = const_cast<float*>(state) + ((i / HW) * 3 + k) * HW + i % HW;
  if (to_state) *s = buf[i]; else buf[i] = *s;
}
__global__ void spi_v_kernel(const float* __restrict__ st_n, float* __restrict__ v, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::spi_v_elem(i, st_n, v, HW);
}
__global__ void spi_step_bwd_kernel(const float* __restrict__ st_i, const float* __restrict__ x0, const float* __restrict__ K,
                                    int64_t K_stride, const float* __restrict__ mu, const float* __restrict__ gv,
                                    float* __restrict__ GX, float* __restrict__ GZ, float* __restrict__ GU,
                                    float* __restrict__ term, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::spi_step_elem(i, st_i, x0, K, K_stride, mu, gv, GX, GZ, GU, term, HW);
}
// out[b * stride] = sum_p term[b, p]; one CTA per image
__global__ void __launch_bounds__(256)
image_sum_kernel(const float* __restrict__ term, float* __restrict__ out, int64_t stride, int HW) {
  __shared__ float red[256];
  const float* t = term + (size_t)blockIdx.x * HW;
  float s = 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) s += t[p];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x * stride] = red[0];
}
// hyper-parameters transposed: P[i*B + b] = sigma_d[b, i], P[(iters + i)*B + b] = mu[b, i]
__global__ void spi_gather_params(const float* __restrict__ sg, const float* __restrict__ mu, int64_t rs, int64_t cs,
                                  float* __restrict__ P, int B, int iters) {
  const int n = B * iters;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = t / B, b = t % B;
    P[t] = sg[b * rs + i * cs];
    P[n + t] = mu[b * rs + i * cs];
  }
}

struct SpiGradOps {
  Denoiser* den; int B, H, W; cudaStream_t st;
  const float* x0; const float* K; int64_t K_stride;
  static constexpr int T = 256;
  int HW() const { return H * W; }
  size_t n() const { return (size_t)B * H * W; }
  unsigned nb() const { return (unsigned)((n() + T - 1) / T); }
  int slot_get(const float* state, float* buf, int k) {
    spi_slot_copy<<<nb(), T, 0, st>>>(state, buf, k, HW(), n(), 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int slot_put(float* state, float* buf, int k) {
    spi_slot_copy<<<nb(), T, 0, st>>>(state, buf, k, HW(), n(), 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int make_v(const float* st_n, float* v) {
    spi_v_kernel<<<nb(), T, 0, st>>>(st_n, v, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gx, float* gv, float* gsig, int64_t stride) {
    return den->vjp(v, sg_i, 1, gx, gv, gsig, stride, B, H, W, st);
  }
  int step(const float* st_i, const float* mu_i, const float* gv, float* gx, float* gz, float* gu, float* term) {
    spi_step_bwd_kernel<<<nb(), T, 0, st>>>(st_i, x0, K, K_stride, mu_i, gv, gx, gz, gu, term, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    image_sum_kernel<<<B, 256, 0, st>>>(term, out, stride, HW());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
};

}  // namespace

int spi_backward(Denoiser* den, const float* states, const float* x0, const float* K, int64_t K_stride, const float* sigma_d,
                 const float* mu, int64_t rs, int64_t cs, int B, int H, int W, int iters, const float* grad_out, float* g_sigma,
                 float* g_mu, float* g_state_in, cudaStream_t st) {
  const size_t n = (size_t)B * H * W;
  PoolBuf bufs[6], P;
  auto body = [&]() -> int {
    for (PoolBuf& b : bufs) TFPNP_TRY(b.alloc(n * sizeof(float), st));
    TFPNP_TRY(P.alloc((size_t)B * iters * 2 * sizeof(float), st));
    spi_gather_params<<<cdiv(B * iters, 256), 256, 0, st>>>(sigma_d, mu, rs, cs, P.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    SpiGradOps ops{den, B, H, W, st, x0, K, K_stride};
    grad_elem::SpiGradBufs w{bufs[0].as<float>(), bufs[1].as<float>(), bufs[2].as<float>(), bufs[3].as<float>(),
                             bufs[4].as<float>(), bufs[5].as<float>()};
    TFPNP_TRY(grad_elem::spi_backward_sequence(ops, states, P.as<float>(), B, H * W, iters, grad_out, g_sigma, g_mu, g_state_in, w));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  };
  const int rc = body();
  for (PoolBuf& b : bufs) b.release();
  P.release();
  return rc;
}

}  // namespace tfpnp

extern "C" int tfpnp_spi_admm_backward(void* denoiser, const float* states, const float* x0, const float* K, int64_t K_stride,
                                       const float* sigma_d, const float* mu, int64_t row_stride, int64_t col_stride, int B,
                                       int H, int W, int iters, const float* grad_out, float* grad_sigma_d, float* grad_mu,
                                       float* grad_state_in, void* stream) {
  using namespace tfpnp;
  TFPNP_CHECK(denoiser && states && x0 && K && sigma_d && mu && grad_out && grad_sigma_d && grad_mu && B > 0 && iters > 0,
              "bad argument");
  g_launch_count = 0;
  return spi_backward(static_cast<Denoiser*>(denoiser), states, x0, K, K_stride, sigma_d, mu, row_stride, col_stride, B, H, W,
                      iters, grad_out, grad_sigma_d, grad_mu, grad_state_in, static_cast<cudaStream_t>(stream));
}
