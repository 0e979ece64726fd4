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
#include <cstdlib>
#include <vector>

namespace tfpnp {
namespace {

// Both orientations of the image are copied into zero-padded scratch (pitch N + 4, data in columns 2..N+1, columns 0, 1,
// N+2, N+3 zero -- written once by CtGeom::reserve, never touched again):
//   P [b][y][2 + x] = img[b][y][x]     rows for the row-driven views
//   PT[b][x][2 + y] = img[b][y][x]     the TRANSPOSE for the column-driven views: consecutive detector bins (lanes) touch
//                                      consecutive addresses (the uncoalesced walk made the projector L1-wavefront bound)
// With the pads a clamped index replaces the four bounds tests of each interpolation (round 2: the projector was
// issue-bound, 34 instructions per ray step -- profiles/r02_ncu_upd_ct.txt).
constexpr int kCtPad = 2;

__global__ void __launch_bounds__(256)
pad_transpose_kernel(const float* __restrict__ img, float* __restrict__ P, float* __restrict__ PT, int N) {
  __shared__ float tile[32][33];
  const int pitch = N + 2 * kCtPad;
  const size_t base = (size_t)blockIdx.z * N * N, pbase = (size_t)blockIdx.z * N * pitch + kCtPad;
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int x = x0 + tx, y = y0 + ty + k;
    const float v = (x < N && y < N) ? img[base + (size_t)y * N + x] : 0.f;
    tile[ty + k][tx] = v;
    if (x < N && y < N) P[pbase + (size_t)y * pitch + x] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int x = y0 + tx, y = x0 + ty + k;                  // transposed coordinates
    if (x < N && y < N) PT[pbase + (size_t)y * pitch + x] = tile[tx][ty + k];
  }
}

// sino[b,v,d] = sum over the driving axis of the linearly interpolated image / max(|cos|,|sin|)  (- y0[b,v,d] if y0)
// Both branches walk `line` = a padded row of P (row-driven views) or PT (column-driven views).  Steps whose two taps
// both fall outside the image add exactly 0, so the walk is cut to the steps where the ray is inside (+- 2 of margin).
// One CTA = 32 rays x kRaySplit warps: warp q takes the steps k_lo + q, k_lo + q + kRaySplit, ... of its 32 rays and the
// partial sums meet in shared memory (one thread per ray walked 256 dependent steps with ~26 warps per SM: latency-bound).
template <int kRaySplit, int UNROLL>
__global__ void __launch_bounds__(32 * kRaySplit)
radon_fwd_kernel(const float* __restrict__ P, const float* __restrict__ PT, const float* __restrict__ y0,
                 float* __restrict__ sino, const float* __restrict__ cs, const float* __restrict__ sn, int N, int D) {
  __shared__ float part[kRaySplit][32];
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int d = blockIdx.x * 32 + lane;
  const int v = blockIdx.y, b = blockIdx.z, V = gridDim.y;
  const float co = cs[v], si = sn[v];
  const float c = (N - 1) * 0.5f;
  const float s = (float)d - (D - 1) * 0.5f;
  const bool col_drive = fabsf(si) >= fabsf(co);
  const float m = col_drive ? fabsf(si) : fabsf(co);
  // position along the interpolated axis at driving index k:  r(k) = (s - (k - c) * a) / bq + c
  const float a = col_drive ? co : si, bq = col_drive ? si : co;
  const float inv_b = 1.0f / bq;
  const int pitch = N + 2 * kCtPad;
  const float* src = (col_drive ? PT : P) + (size_t)b * N * pitch + kCtPad;
  // driving indices with r(k) in [-1, N]:  k(r) = c + (s - (r - c) bq) / a
  int k_lo = 0, k_hi = N - 1;
  if (a != 0.f) {
    const float ka = c + (s - (-1.f - c) * bq) / a, kb = c + (s - ((float)N - c) * bq) / a;
    const float lo = fminf(fmaxf(fminf(ka, kb), -2.f), (float)N + 1.f);
    const float hi = fminf(fmaxf(fmaxf(ka, kb), -2.f), (float)N + 1.f);
    k_lo = max(0, (int)floorf(lo) - 2);
    k_hi = min(N - 1, (int)ceilf(hi) + 2);
  }
  if (d >= D) k_hi = -1;
  float acc = 0.f;
  float t = (float)(k_lo + q) - c;                // (float)k - c exactly: both are multiples of 0.5 far below 2^24
  const float* line = src + (size_t)(k_lo + q) * pitch;
#pragma unroll UNROLL
  for (int k = k_lo + q; k <= k_hi; k += kRaySplit) {
    const float r = __fadd_rn(__fmul_rn(__fsub_rn(s, __fmul_rn(t, a)), inv_b), c);
    const float fl = floorf(r);
    const float f = r - fl;
    const int i0 = min(max((int)fl, -kCtPad), N);  // clamped into the pads: both taps read 0 there
    acc += (1.f - f) * line[i0] + f * line[i0 + 1];
    t += (float)kRaySplit;
    line += (size_t)kRaySplit * pitch;
  }
  part[q][lane] = acc;
  __syncthreads();
  if (q == 0 && d < D) {
    float tot = part[0][lane];
#pragma unroll
    for (int i = 1; i < kRaySplit; ++i) tot += part[i][lane];
    const size_t o = ((size_t)b * V + v) * D + d;
    const float r = tot / m;
    sino[o] = y0 ? r - y0[o] : r;
  }
}

// per-view constants in shared memory: (cos, sin, 1 / max(|cos|, |sin|))
__device__ __forceinline__ void load_view_table(float4* tab, const float* __restrict__ cs, const float* __restrict__ sn, int V) {
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float co = cs[v], si = sn[v];
    tab[v] = make_float4(co, si, 1.0f / fmaxf(fabsf(si), fabsf(co)), 0.f);   // one division per view
  }
  __syncthreads();
}

// GUARD = false: the caller guarantees 0 <= floor(d*) <= D - 2 for every pixel (unit-norm (cos, sin) and
// D >= sqrt(2) N -- CtGeom::bins_always_inside); the clamp only keeps a bad table memory-safe.
template <bool GUARD>
__device__ __forceinline__ float backproject_pixel(const float* __restrict__ sg, const float4* tab, int V, int D, float xx,
                                                   float yy) {
  float acc = 0.f;
  const float half = (D - 1) * 0.5f;
#pragma unroll 4
  for (int v = 0; v < V; ++v) {
    const float4 tv = tab[v];
    const float inv_m = tv.z;
    const float dstar = __fadd_rn(__fadd_rn(__fmul_rn(xx, tv.x), __fmul_rn(yy, tv.y)), half);
    const float fl = floorf(dstar);
    const float* row = sg + (size_t)v * D;
    const float w0 = fmaxf(1.f - fabsf(fl - dstar) * inv_m, 0.f) * inv_m;
    const float w1 = fmaxf(1.f - fabsf((fl + 1.f) - dstar) * inv_m, 0.f) * inv_m;
    if (GUARD) {
      const int d0 = (int)fl;
      if (d0 >= 0 && d0 < D) acc += row[d0] * w0;
      if (d0 + 1 >= 0 && d0 + 1 < D) acc += row[d0 + 1] * w1;
    } else {
      const int d0 = min(max((int)fl, 0), D - 2);
      acc += row[d0] * w0;
      acc += row[d0 + 1] * w1;
    }
  }
  return acc;
}

template <bool GUARD>
__global__ void __launch_bounds__(256)
radon_bwd_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ cs,
                 const float* __restrict__ sn, int N, int V, int D) {
  extern __shared__ float4 tab[];
  load_view_table(tab, cs, sn, V);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= N * N) return;
  const float c = (N - 1) * 0.5f;
  float xx = (float)(p % N) - c, yy = (float)(p / N) - c;
  img[(size_t)b * N * N + p] = backproject_pixel<GUARD>(sino + (size_t)b * V * D, tab, V, D, xx, yy);
}

template <bool GUARD>
__global__ void __launch_bounds__(256)
ct_bwd_update_kernel(const float* __restrict__ resid, const float* __restrict__ x, float* __restrict__ z,
                     float* __restrict__ u, float* __restrict__ d, const float* __restrict__ cs,
                     const float* __restrict__ sn, const float* __restrict__ mu, const float* __restrict__ tau,
                     float inv_opnorm2, int N, int V, int D) {
  extern __shared__ float4 tab[];
  load_view_table(tab, cs, sn, V);
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (p >= N * N) return;
  const float c = (N - 1) * 0.5f;
  float xx = (float)(p % N) - c, yy = (float)(p / N) - c;
  float bp = backproject_pixel<GUARD>(resid + (size_t)b * V * D, tab, V, D, xx, yy) * inv_opnorm2;
  size_t i = (size_t)b * N * N + p;
  float zz = z[i], uu = u[i], xv = x[i];
  zz = zz - tau[b] * (bp + mu[b] * (zz - (xv + uu)));   // solver.py:46
  uu = uu + xv - zz;                                     // solver.py:49
  z[i] = zz; u[i] = uu; d[i] = zz - uu;                  // next denoiser input z - u (solver.py:39)
}

// Windowed back-projection for a 16 x 16 pixel tile (requires bins_always_inside): per view the tile's pixels project into
// fewer than 24 consecutive bins, so the CTA first copies a 32-bin window of every view into shared memory (all loads
// independent: one L2 round trip) and the per-pixel gather then reads shared memory only.  (The row-per-CTA gather from
// global memory touched ~62 KB of sinogram per CTA and waited on L2: long-scoreboard 17 warps per issue slot.)
constexpr int kBpWin = 32;

struct BpSmem {
  float4* tab;    // [V] cos, sin, 1/max(|cos|,|sin|), window start (as float)
  float* win;     // [V][kBpWin]
};
__device__ __forceinline__ BpSmem bp_smem(int V) {
  extern __shared__ float4 bp_raw[];
  return {bp_raw, reinterpret_cast<float*>(bp_raw + V)};
}
inline size_t bp_smem_bytes(int V) { return (size_t)V * (sizeof(float4) + kBpWin * sizeof(float)); }

__device__ __forceinline__ void bp_stage(const BpSmem& sm, const float* __restrict__ sg, const float* __restrict__ cs,
                                         const float* __restrict__ sn, int V, int D, float x0, float y0) {
  const float half = (D - 1) * 0.5f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float co = cs[v], si = sn[v];
    // smallest d* over the tile's corners, one bin of margin for the different rounding of the per-pixel d*
    const float e0 = x0 * co + y0 * si, e1 = (x0 + 15.f) * co + y0 * si, e2 = x0 * co + (y0 + 15.f) * si,
                e3 = (x0 + 15.f) * co + (y0 + 15.f) * si;
    const float dmin = fminf(fminf(e0, e1), fminf(e2, e3)) + half;
    const int start = min(max((int)floorf(dmin) - 1, 0), D - kBpWin);
    sm.tab[v] = make_float4(co, si, 1.0f / fmaxf(fabsf(si), fabsf(co)), (float)start);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < V * kBpWin; i += blockDim.x) {
    const int v = i / kBpWin, k = i % kBpWin;
    sm.win[i] = sg[(size_t)v * D + (int)sm.tab[v].w + k];
  }
  __syncthreads();
}

__device__ __forceinline__ float bp_pixel(const BpSmem& sm, int V, int D, float xx, float yy) {
  float acc = 0.f;
  const float half = (D - 1) * 0.5f;
#pragma unroll 4
  for (int v = 0; v < V; ++v) {
    const float4 tv = sm.tab[v];
    const float inv_m = tv.z;
    const float dstar = __fadd_rn(__fadd_rn(__fmul_rn(xx, tv.x), __fmul_rn(yy, tv.y)), half);
    const float fl = floorf(dstar);
    const float w0 = fmaxf(1.f - fabsf(fl - dstar) * inv_m, 0.f) * inv_m;
    const float w1 = fmaxf(1.f - fabsf((fl + 1.f) - dstar) * inv_m, 0.f) * inv_m;
    const int o = min(max((int)(fl - tv.w), 0), kBpWin - 2);   // (the clamp only keeps a bad table memory-safe)
    const float* row = sm.win + v * kBpWin + o;
    acc += row[0] * w0;
    acc += row[1] * w1;
  }
  return acc;
}

__global__ void __launch_bounds__(256)
radon_bwd_win_kernel(const float* __restrict__ sino, float* __restrict__ img, const float* __restrict__ cs,
                     const float* __restrict__ sn, int N, int V, int D) {
  const BpSmem sm = bp_smem(V);
  const int b = blockIdx.z;
  const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
  const float c = (N - 1) * 0.5f;
  bp_stage(sm, sino + (size_t)b * V * D, cs, sn, V, D, (float)(blockIdx.x * 16) - c, (float)(blockIdx.y * 16) - c);
  img[(size_t)b * N * N + (size_t)py * N + px] = bp_pixel(sm, V, D, (float)px - c, (float)py - c);
}

__global__ void __launch_bounds__(256)
ct_bwd_update_win_kernel(const float* __restrict__ resid, const float* __restrict__ x, float* __restrict__ z,
                         float* __restrict__ u, float* __restrict__ d, const float* __restrict__ cs,
                         const float* __restrict__ sn, const float* __restrict__ mu, const float* __restrict__ tau,
                         float inv_opnorm2, int N, int V, int D) {
  const BpSmem sm = bp_smem(V);
  const int b = blockIdx.z;
  const int px = blockIdx.x * 16 + (threadIdx.x & 15), py = blockIdx.y * 16 + (threadIdx.x >> 4);
  const float c = (N - 1) * 0.5f;
  const size_t i = (size_t)b * N * N + (size_t)py * N + px;
  float zz = z[i], uu = u[i], xv = x[i];                 // in flight while the windows are staged
  bp_stage(sm, resid + (size_t)b * V * D, cs, sn, V, D, (float)(blockIdx.x * 16) - c, (float)(blockIdx.y * 16) - c);
  const float bp = bp_pixel(sm, V, D, (float)px - c, (float)py - c) * inv_opnorm2;
  zz = zz - tau[b] * (bp + mu[b] * (zz - (xv + uu)));   // solver.py:46
  uu = uu + xv - zz;                                     // solver.py:49
  z[i] = zz; u[i] = uu; d[i] = zz - uu;                  // next denoiser input z - u (solver.py:39)
}

// the windowed kernels need whole 16 x 16 tiles, provably-inside bins, >= 32 bins and the windows in 48 KB
bool bp_windowed(const CtGeom& g) {
  return g.bins_always_inside && g.N % 16 == 0 && g.det >= kBpWin && bp_smem_bytes(g.views) <= 48 * 1024;
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
  // |x cos + y sin| <= |(x, y)| needs unit-norm rows; then every pixel centre projects inside [0.2, D - 1.2] when D >= sqrt(2) N
  bins_always_inside = det >= 2 && (double)det >= std::sqrt(2.0) * N;
  for (int i = 0; i < views; ++i)
    if (std::fabs((double)cos_host[i] * cos_host[i] + (double)sin_host[i] * sin_host[i] - 1.0) > 1e-4) bins_always_inside = false;
  TFPNP_TRY(cs.alloc(views * sizeof(float)));
  TFPNP_TRY(sn.alloc(views * sizeof(float)));
  TFPNP_CUDA_OK(cudaMemcpy(cs.p, cos_host, views * sizeof(float), cudaMemcpyHostToDevice));
  TFPNP_CUDA_OK(cudaMemcpy(sn.p, sin_host, views * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}

size_t CtGeom::scratch_bytes(int B) const { return (size_t)2 * B * N * (N + 2 * kCtPad) * sizeof(float); }

int CtGeom::reserve(int B) const {
  if (tbuf.bytes >= scratch_bytes(B)) return 0;
  TFPNP_TRY(tbuf.alloc(scratch_bytes(B)));
  TFPNP_CUDA_OK(cudaMemset(tbuf.p, 0, tbuf.bytes));   // the pad columns stay zero for the life of the buffer
  TFPNP_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

int radon_forward(const CtGeom& g, const float* img, const float* y0, float* sino, int B, cudaStream_t st) {
  TFPNP_CHECK(g.tbuf.bytes >= g.scratch_bytes(B), "CtGeom::reserve(%d) not called", B);
  float* P = g.tbuf.as<float>();
  float* PT = P + (size_t)B * g.N * (g.N + 2 * kCtPad);
  pad_transpose_kernel<<<dim3(cdiv(g.N, 32), cdiv(g.N, 32), B), 256, 0, st>>>(img, P, PT, g.N);
  TFPNP_COUNT_LAUNCH();
  // (4 warps per ray group, unroll 4: 81 us per iteration at 8 x 256^2 x 60 views; 4/8, 8/4, 8/8, 2/8 measured 83-88)
  radon_fwd_kernel<4, 4><<<dim3(cdiv(g.det, 32), g.views, B), 32 * 4, 0, st>>>(P, PT, y0, sino, g.cs.as<float>(),
                                                                              g.sn.as<float>(), g.N, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int radon_backward(const CtGeom& g, const float* sino, float* img, int B, cudaStream_t st) {
  if (bp_windowed(g)) {
    radon_bwd_win_kernel<<<dim3(g.N / 16, g.N / 16, B), 256, bp_smem_bytes(g.views), st>>>(sino, img, g.cs.as<float>(),
                                                                                           g.sn.as<float>(), g.N, g.views, g.det);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const dim3 grid(cdiv(g.N * g.N, 256), B);
  const size_t sm = (size_t)g.views * sizeof(float4);
  TFPNP_CHECK(sm <= 48 * 1024, "ct: %d views exceed the per-view table in shared memory (3072)", g.views);
  if (g.bins_always_inside)
    radon_bwd_kernel<false><<<grid, 256, sm, st>>>(sino, img, g.cs.as<float>(), g.sn.as<float>(), g.N, g.views, g.det);
  else
    radon_bwd_kernel<true><<<grid, 256, sm, st>>>(sino, img, g.cs.as<float>(), g.sn.as<float>(), g.N, g.views, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int ct_update(const CtGeom& g, const float* x, float* z, float* u, float* d, float* resid, const float* y0,
              float inv_opnorm2, const float* mu, const float* tau, int B, cudaStream_t st) {
  TFPNP_TRY(radon_forward(g, z, y0, resid, B, st));
  if (bp_windowed(g)) {
    ct_bwd_update_win_kernel<<<dim3(g.N / 16, g.N / 16, B), 256, bp_smem_bytes(g.views), st>>>(
        resid, x, z, u, d, g.cs.as<float>(), g.sn.as<float>(), mu, tau, inv_opnorm2, g.N, g.views, g.det);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  const dim3 grid(cdiv(g.N * g.N, 256), B);
  const size_t sm = (size_t)g.views * sizeof(float4);
  TFPNP_CHECK(sm <= 48 * 1024, "ct: %d views exceed the per-view table in shared memory (3072)", g.views);
  if (g.bins_always_inside)
    ct_bwd_update_kernel<false><<<grid, 256, sm, st>>>(resid, x, z, u, d, g.cs.as<float>(), g.sn.as<float>(), mu, tau,
                                                        inv_opnorm2, g.N, g.views, g.det);
  else
    ct_bwd_update_kernel<true><<<grid, 256, sm, st>>>(resid, x, z, u, d, g.cs.as<float>(), g.sn.as<float>(), mu, tau,
                                                       inv_opnorm2, g.N, g.views, g.det);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace tfpnp
