// Sparse-view CT: parallel-beam Radon pair and the inexact-ADMM z-step fused with the dual
// update (tasks/ct/solver.py:39-49):
//     z -= tau ( A^T(A z - y0) / opnorm^2 + mu (z - (x + u)) );  u += x - z;  d = z - u
//
// The reference delegates A / A^T to the third-party `torch_radon` (absent, unpinned), so the
// discretisation is this build's own, on the geometry the reference fixes
// (tfpnp/utils/transforms.py:487-491): angles linspace(0, 179pi/180, V), det_count =
// ceil(sqrt(2) N), unit spacing, rotation about the image centre, no circle clipping.
//   A   : Joseph-type ray-driven projector (step over the driving axis, linear interpolation
//         on the other, weight 1/max(|cos|,|sin|));
//   A^T : its exact transpose written as a gather (each pixel reads <= 2 detector bins per view),
// so both directions are atomic-free gathers; parity is against oracle/pnp_oracle.py's
// restatement of the same formulas (PARITY UNPINNED w.r.t. torch_radon).
// Three launches per iteration: transpose (for the column-driven views), forward (+ "- y0") and backprojection
// fused with the update.
#include "tasks.cuh"
#include <cmath>
#include <vector>

namespace tfpnp {
namespace {

// imgT[b][j][i] = img[b][i][j]: the column-driven views walk the image column by column, so reading the TRANSPOSED
// image makes consecutive detector bins (consecutive lanes) touch consecutive addresses (measured: the uncoalesced
// walk made the projector L1-wavefront bound, 170 us for 8 x 256^2 x 60 views)
__global__ void __launch_bounds__(256)
transpose_kernel(const float* __restrict__ img, float* __restrict__ imgT, int N) {
  __shared__ float tile[32][33];
  const size_t base = (size_t)blockIdx.z * N * N;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int x = x0 + tx, y = y0 + ty + k;
    tile[ty + k][tx] = (x < N && y < N) ? img[base + (size_t)y * N + x] : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int x = y0 + tx, y = x0 + ty + k;                  // transposed coordinates
    if (x < N && y < N) imgT[base + (size_t)y * N + x] = tile[tx][ty + k];
  }
}

// sino[b,v,d] = sum over the driving axis of the linearly interpolated image / max(|cos|,|sin|)  (- y0[b,v,d] if y0)
// Both branches walk `line` = a row of `src` (the image for row-driven views, its transpose for column-driven ones).
__global__ void __launch_bounds__(128)
radon_fwd_kernel(const float* __restrict__ img, const float* __restrict__ imgT, const float* __restrict__ y0,
                 float* __restrict__ sino, const float* __restrict__ cs, const float* __restrict__ sn, int N, int D) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = blockIdx.y, b = blockIdx.z, V = gridDim.y;
  if (d >= D) return;
  const float co = cs[v], si = sn[v];
  const float c = (N - 1) * 0.5f;
  const float s = (float)d - (D - 1) * 0.5f;
  const bool col_drive = fabsf(si) >= fabsf(co);
  const float m = col_drive ? fabsf(si) : fabsf(co);
  // position along the interpolated axis at driving index k:  r(k) = (s - (k - c) * a) / bq + c
  const float a = col_drive ? co : si, bq = col_drive ? si : co;
  const float inv_b = 1.0f / bq;
  const float* src = (col_drive ? imgT : img) + (size_t)b * N * N;
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < N; ++k) {
    const float t = (float)k - c;
    const float r = __fadd_rn(__fmul_rn(__fsub_rn(s, __fmul_rn(t, a)), inv_b), c);
    const float fl = floorf(r);
    const float f = r - fl;
    const int i0 = (int)fl;
    const float* line = src + (size_t)k * N;
    const float v0 = (i0 >= 0 && i0 < N) ? line[i0] : 0.f;
    const float v1 = (i0 + 1 >= 0 && i0 + 1 < N) ? line[i0 + 1] : 0.f;
    acc += (1.f - f) * v0 + f * v1;
  }
  const size_t o = ((size_t)b * V + v) * D + d;
  const float r = acc / m;
  sino[o] = y0 ? r - y0[o] : r;
}

__device__ __forceinline__ float backproject_pixel(const float* __restrict__ sg, const float* __restrict__ cs,
                                                   const float* __restrict__ sn, int V, int D, float xx,
                                                   float yy) {
  float acc = 0.f;
  const float half = (D - 1) * 0.5f;
  for (int v = 0; v < V; ++v) {
    const float co = cs[v], si = sn[v];
    const float inv_m = 1.0f / fmaxf(fabsf(si), fabsf(co));     // one division per view instead of four
    float dstar = __fadd_rn(__fadd_rn(__fmul_rn(xx, co), __fmul_rn(yy, si)), half);
    float fl = floorf(dstar);
    int d0 = (int)fl;
    const float* row = sg + (size_t)v * D;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      int dd = d0 + k;
      float w = fmaxf(1.f - fabsf((fl + (float)k) - dstar) * inv_m, 0.f) * inv_m;
      if (dd >= 0 && dd < D) acc += row[dd] * w;
    }
  }
  return acc;
}

__global__ void __launch_bounds__(256)
radon_bwd_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ cs,
                 const float* __restrict__ sn, int N, int V, int D) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= N * N) return;
  const float c = (N - 1) * 0.5f;
  float xx = (float)(p % N) - c, yy = (float)(p / N) - c;
  img[(size_t)b * N * N + p] = backproject_pixel(sino + (size_t)b * V * D, cs, sn, V, D, xx, yy);
}

__global__ void __launch_bounds__(256)
ct_bwd_update_kernel(const float* __restrict__ resid, const float* __restrict__ x, float* __restrict__ z,
                     float* __restrict__ u, float* __restrict__ d, const float* __restrict__ cs,
                     const float* __restrict__ sn, const float* __restrict__ mu, const float* __restrict__ tau,
                     float inv_opnorm2, int N, int V, int D) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= N * N) return;
  const float c = (N - 1) * 0.5f;
  float xx = (float)(p % N) - c, yy = (float)(p / N) - c;
  float bp = backproject_pixel(resid + (size_t)b * V * D, cs, sn, V, D, xx, yy) * inv_opnorm2;
  size_t i = (size_t)b * N * N + p;
  float zz = z[i], uu = u[i], xv = x[i];
  zz = zz - tau[b] * (bp + mu[b] * (zz - (xv + uu)));   // solver.py:46
  uu = uu + xv - zz;                                     // solver.py:49
  z[i] = zz; u[i] = uu; d[i] = zz - uu;                  // next denoiser input z - u (solver.py:39)
}

}  // namespace

int CtGeom::init(int N_, int views_) {
  N = N_; views = views_;
  det = (int)std::ceil(std::sqrt(2.0) * N);              // transforms.py:489
  // angles = torch.linspace(0, 179/180*pi, views) in fp32 (transforms.py:488)
  std::vector<float> c(views), s(views);
  const float end = (float)(179.0 / 180.0 * M_PI);
  const float step = views > 1 ? end / (float)(views - 1) : 0.f;
  for (int i = 0; i < views; ++i) {
    float a = (i < views / 2) ? step * (float)i : end - step * (float)(views - i - 1);
    c[i] = (float)std::cos((double)a);
    s[i] = (float)std::sin((double)a);
  }
  return set_tables(c.data(), s.data());
}

int CtGeom::set_tables(const float* cos_host, const float* sin_host) {
  TFPNP_TRY(cs.alloc(views * sizeof(float)));
  TFPNP_TRY(sn.alloc(views * sizeof(float)));
  TFPNP_CUDA_OK(cudaMemcpy(cs.p, cos_host, views * sizeof(float), cudaMemcpyHostToDevice));
  TFPNP_CUDA_OK(cudaMemcpy(sn.p, sin_host, views * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

int CtGeom::reserve(int B) const {
  return tbuf.alloc((size_t)B * N * N * sizeof(float));
}

int radon_forward(const CtGeom& g, const float* img, const float* y0, float* sino, int B, cudaStream_t st) {
  TFPNP_CHECK(g.tbuf.bytes >= (size_t)B * g.N * g.N * sizeof(float), "CtGeom::reserve(%d) not called", B);
  transpose_kernel<<<dim3(cdiv(g.N, 32), cdiv(g.N, 32), B), 256, 0, st>>>(img, g.tbuf.as<float>(), g.N);
  TFPNP_COUNT_LAUNCH();
  radon_fwd_kernel<<<dim3(cdiv(g.det, 128), g.views, B), 128, 0, st>>>(img, g.tbuf.as<float>(), y0, sino, g.cs.as<float>(),
                                                                        g.sn.as<float>(), g.N, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int radon_backward(const CtGeom& g, const float* sino, float* img, int B, cudaStream_t st) {
  radon_bwd_kernel<<<dim3(cdiv(g.N * g.N, 256), B), 256, 0, st>>>(sino, img, g.cs.as<float>(), g.sn.as<float>(),
                                                                    g.N, g.views, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int ct_update(const CtGeom& g, const float* x, float* z, float* u, float* d, float* resid, const float* y0,
              float inv_opnorm2, const float* mu, const float* tau, int B, cudaStream_t st) {
  TFPNP_TRY(radon_forward(g, z, y0, resid, B, st));
  ct_bwd_update_kernel<<<dim3(cdiv(g.N * g.N, 256), B), 256, 0, st>>>(
      resid, x, z, u, d, g.cs.as<float>(), g.sn.as<float>(), mu, tau, inv_opnorm2, g.N, g.views, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tfpnp
