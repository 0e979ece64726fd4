// CS-MRI data-fidelity step fused with the ADMM primal/dual update
// (tasks/csmri/solver.py:47-55):
//     Z = fft2c(x + u);  Z[mask] = (mu Z + y0)[mask] / (1 + mu);  z = ifft2c(Z);  u += x - z
// plus the next denoiser input d = Re(z - u).
//
// The centred transforms (transforms.py:68-103) are computed as plain FFTs: for even N,
// fft2c(x)[k + N/2] = (-1)^(k1+k2) FFT(x)[k] and the two sign/roll pairs cancel around the
// pointwise step once y0 is pre-multiplied by (-1)^(k1+k2) and y0/mask are pre-rolled
// (csmri_prep, once per solver call).  Sign flips and the 1/N scale (a power of two) are
// exact, so the masked update is the reference's arithmetic on the same operands.
//
// Three launches per iteration, all HBM/L2-streaming with coalesced float2 access:
//   rows_fwd : warp per image row   (x+u) -> row FFT -> T
//   cols     : 16 columns per CTA through a padded smem tile: col FFT -> DC -> inverse col FFT
//   rows_inv : warp per image row   inverse row FFT -> z, u, d
#include "tasks.cuh"
#include "fft.cuh"
#include "sm100.cuh"
#include <cstdlib>

namespace tfpnp {
namespace {

constexpr int ROWS_PER_CTA = 8;   // warps per CTA in the row kernels
constexpr int COLS_PER_CTA = 8;   // columns (= warps) per CTA in the column kernel: more, smaller CTAs per SM overlap the load / FFT / store phases

__global__ void csmri_prep_kernel(const float2* __restrict__ y0, const uint8_t* __restrict__ mask,
                                  float2* __restrict__ y0p, uint8_t* __restrict__ maskp, int N, int R) {
  // one thread per (b, c, r); r fastest so the transposed write is coalesced
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  int r = i % N, c = (i / N) % N;
  size_t b = i / ((size_t)N * N);
  int kx = fft_pos_to_freq(c, R), ky = fft_pos_to_freq(r, R);
  int sx = (kx + N / 2) % N, sy = (ky + N / 2) % N;      // position in the centred spectrum
  size_t src = (b * N + sy) * N + sx;
  float2 v = y0[src];
  if ((kx + ky) & 1) { v.x = -v.x; v.y = -v.y; }
  y0p[i] = v;
  maskp[i] = mask[src];
}

template <int R>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
csmri_rows_fwd(const float* __restrict__ x, const float2* __restrict__ u, float2* __restrict__ T) {
  constexpr int N = 32 * R;
  sm100::pdl_launch_dependents();   // (launched with programmatic stream serialisation: the twiddle loads overlap the denoiser's tail)
  WarpFFT<R> f;
  f.init();
  size_t row = (size_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  const float* xr = x + row * N;
  const float2* ur = u + row * N;
  float2 v[R];
  sm100::pdl_wait_then(xr, ur);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 uu = ur[32 * j + f.lane];
    v[j] = make_float2(xr[32 * j + f.lane] + uu.x, uu.y);
  }
  f.forward(v);
  float2* tr = T + row * N;
#pragma unroll
  for (int j = 0; j < R; ++j) tr[32 * j + f.lane] = v[j];
}

// (launch bounds: 6 CTAs per SM, so that the 768 CTAs of the 48 x 128^2 case are ONE wave -- at the 4 per SM the register
// count allowed before, the 1.3 waves made this latency-bound kernel take two round trips)
template <int R>
__global__ void __launch_bounds__(COLS_PER_CTA * 32, R <= 4 ? 6 : 3)
csmri_cols(float2* __restrict__ T, const float2* __restrict__ y0p, const uint8_t* __restrict__ maskp,
           const float* __restrict__ mu) {
  constexpr int N = 32 * R;
  constexpr int PITCH = COLS_PER_CTA + 1;  // float2 pitch 17 -> conflict-free column walks
  __shared__ float2 tile[N * PITCH];
  const int b = blockIdx.y, c0 = blockIdx.x * COLS_PER_CTA;
  float2* Tb = T + (size_t)b * N * N;
  sm100::pdl_launch_dependents();
  WarpFFT<R> f;
  f.init();
  sm100::pdl_wait_then(Tb);
  for (int i = threadIdx.x; i < N * COLS_PER_CTA; i += COLS_PER_CTA * 32) {
    int r = i / COLS_PER_CTA, cc = i % COLS_PER_CTA;
    tile[r * PITCH + cc] = Tb[(size_t)r * N + c0 + cc];
  }
  __syncthreads();
  const int w = threadIdx.x >> 5;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = tile[(32 * j + f.lane) * PITCH + w];
  f.forward(v);
  const float m = mu[b];
  const float inv_n = 1.0f / (float)N;  // ortho 2-D scale, exact power of two
  const size_t col = ((size_t)b * N + c0 + w) * N;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 zf = make_float2(v[j].x * inv_n, v[j].y * inv_n);
    if (maskp[col + 32 * j + f.lane]) {        // z[mask] = ((mu z + y0)/(1+mu))[mask], solver.py:50-51
      float2 y = y0p[col + 32 * j + f.lane];
      zf.x = (m * zf.x + y.x) / (1.0f + m);
      zf.y = (m * zf.y + y.y) / (1.0f + m);
    }
    v[j] = zf;
  }
  f.inverse(v);
#pragma unroll
  for (int j = 0; j < R; ++j) tile[(32 * j + f.lane) * PITCH + w] = v[j];
  __syncthreads();
  for (int i = threadIdx.x; i < N * COLS_PER_CTA; i += COLS_PER_CTA * 32) {
    int r = i / COLS_PER_CTA, cc = i % COLS_PER_CTA;
    Tb[(size_t)r * N + c0 + cc] = tile[r * PITCH + cc];
  }
}

template <int R>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
csmri_rows_inv(const float2* __restrict__ T, const float* __restrict__ x, float2* __restrict__ z,
               float2* __restrict__ u, float* __restrict__ d) {
  constexpr int N = 32 * R;
  sm100::pdl_launch_dependents();
  WarpFFT<R> f;
  f.init();
  size_t row = (size_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
  const float2* tr = T + row * N;
  float2 v[R];
  const float* xp = x;
  float2* up = u;
  sm100::pdl_wait_then(tr, xp, up);
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = tr[32 * j + f.lane];
  f.inverse(v);
  const float inv_n = 1.0f / (float)N;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    size_t i = row * N + 32 * j + f.lane;
    float2 zz = make_float2(v[j].x * inv_n, v[j].y * inv_n);
    float2 uu = up[i];
    float xx = xp[i];
    uu.x = uu.x + xx - zz.x;   // u = u + x - z (solver.py:55), Im(x) = 0
    uu.y = uu.y - zz.y;
    z[i] = zz;
    up[i] = uu;
    d[i] = zz.x - uu.x;        // complex2real(z - u) (solver.py:45)
  }
}

// ---- one launch per iteration: the whole update of an image inside a cluster of CL CTAs ---------------------------
// CTA `rank` of the cluster owns image rows [rank*N/CL, (rank+1)*N/CL) for the two row passes and columns (frequency
// positions) [rank*N/CL, ...) for the column pass.  CL = 4 at 128x128: 192 CTAs with 66 KB of shared memory each fit the
// chip in ONE wave (the first version, CL = 2, was 96 x 2 CTAs of 132 KB = 1.3 waves of one CTA per SM, which is why it
// lost end to end although it saved two launches).  The two transposes between the passes go through shared memory:
// every warp scatters its FFT output into the [col][row] (then [row][col]) tile of the CTA that owns the column (row)
// -- its own tile or, through distributed shared memory (st.shared::cluster), its peer's -- so the intermediate T
// never touches L2/HBM.  Same arithmetic, in the same order, as the three-kernel path above (equal up to the
// compiler's FMA contraction inside the butterflies).
__device__ __forceinline__ void st_cluster_f2(const float2* local_ptr, uint32_t cta, float2 v) {
  uint32_t la = static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr)), ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(cta));
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(v.x), "f"(v.y) : "memory");
}

constexpr int FUSED_WARPS = 16;

template <int R, int CL>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 1)
csmri_fused(const float* __restrict__ x, float2* __restrict__ z, float2* __restrict__ u, float* __restrict__ d,
            const float2* __restrict__ y0p, const uint8_t* __restrict__ maskp, const float* __restrict__ mu) {
  constexpr int N = 32 * R, HALF = N / CL, PITCH = N + 1, RPW = HALF / FUSED_WARPS;   // HALF = rows (columns) per CTA
  static_assert(RPW >= 1, "csmri_fused needs N / CL >= 16");
  extern __shared__ float2 fsm[];
  float2* tA = fsm;                     // [HALF cols][PITCH]: column-major input of the column pass
  float2* tB = fsm + HALF * PITCH;      // [HALF rows][PITCH]: row-major input of the inverse row pass
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // distributed shared memory may only be touched once the peer CTA has started: arrive now, wait before the first
  // remote store (without this the scatter of pass 1 raced with the peer's launch as soon as the prologue got short)
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int b = blockIdx.x / CL;
  const int warp = threadIdx.x >> 5;
  WarpFFT<R> f;
  f.init();                             // twiddles: overlaps the tail of the denoiser's last kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  const size_t img = (size_t)b * N * N;
  float2 v[R];
  // Every pass first issues ALL the global loads of the warp's RPW rows / columns (one latency per pass instead of
  // one per row), then runs the FFTs out of registers.
  // ---- pass 1: (x + u) -> row FFT -> scatter by column owner
  {
    float2 in[RPW][R];
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = rank * HALF + warp * RPW + i;
      const float* xr = x + img + (size_t)r * N;
      const float2* ur = u + img + (size_t)r * N;
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const float2 uu = ur[32 * j + f.lane];
        in[i][j] = make_float2(xr[32 * j + f.lane] + uu.x, uu.y);
      }
    }
#pragma unroll
    for (int i = 0; i < RPW; ++i) {
      const int r = rank * HALF + warp * RPW + i;
#pragma unroll
      for (int j = 0; j < R; ++j) v[j] = in[i][j];
      f.forward(v);
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const int c = 32 * j + f.lane;
        st_cluster_f2(tA + (c % HALF) * PITCH + r, c / HALF, v[j]);
      }
    }
  }
  // data-consistency operands of this warp's columns: constant during the solver call, fetched before the barrier
  float2 yv[RPW][R];
  bool mk[RPW][R];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const size_t col = ((size_t)b * N + rank * HALF + warp * RPW + i) * N;
#pragma unroll
    for (int j = 0; j < R; ++j) {
      mk[i][j] = maskp[col + 32 * j + f.lane] != 0;
      yv[i][j] = y0p[col + 32 * j + f.lane];
    }
  }
  const float m = mu[b];
  const float inv_n = 1.0f / (float)N;
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---- pass 2: column FFT -> masked data consistency -> inverse column FFT -> scatter by row owner
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int cc = warp * RPW + i, c = rank * HALF + cc;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = tA[cc * PITCH + 32 * j + f.lane];
    f.forward(v);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float2 zf = make_float2(v[j].x * inv_n, v[j].y * inv_n);
      if (mk[i][j]) {                            // z[mask] = ((mu z + y0)/(1+mu))[mask], solver.py:50-51
        zf.x = (m * zf.x + yv[i][j].x) / (1.0f + m);
        zf.y = (m * zf.y + yv[i][j].y) / (1.0f + m);
      }
      v[j] = zf;
    }
    f.inverse(v);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const int r = 32 * j + f.lane;
      st_cluster_f2(tB + (r % HALF) * PITCH + c, r / HALF, v[j]);
    }
  }
  // the dual-update operands of this warp's rows (x, u): fetched before the barrier
  float2 uv[RPW][R];
  float xv[RPW][R];
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const size_t row = img + (size_t)(rank * HALF + warp * RPW + i) * N;
#pragma unroll
    for (int j = 0; j < R; ++j) { uv[i][j] = u[row + 32 * j + f.lane]; xv[i][j] = x[row + 32 * j + f.lane]; }
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  // ---- pass 3: inverse row FFT -> z, u += x - z, d = Re(z - u)
#pragma unroll
  for (int i = 0; i < RPW; ++i) {
    const int rr = warp * RPW + i, r = rank * HALF + rr;
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = tB[rr * PITCH + 32 * j + f.lane];
    f.inverse(v);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const size_t idx = img + (size_t)r * N + 32 * j + f.lane;
      const float2 zz = make_float2(v[j].x * inv_n, v[j].y * inv_n);
      float2 uu = uv[i][j];
      uu.x = uu.x + xv[i][j] - zz.x;   // u = u + x - z (solver.py:55), Im(x) = 0
      uu.y = uu.y - zz.y;
      z[idx] = zz;
      u[idx] = uu;
      d[idx] = zz.x - uu.x;            // complex2real(z - u) (solver.py:45)
    }
  }
}

template <int R, int CL>
int launch_fused(const float* x, float2* z, float2* u, float* d, const float2* y0p, const uint8_t* maskp,
                 const float* mu, int B, cudaStream_t st) {
  constexpr int N = 32 * R;
  constexpr int smem = 2 * (N / CL) * (N + 1) * (int)sizeof(float2);
  static unsigned long long attr_set = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_set >> (dev & 63) & 1ull)) {
    TFPNP_CUDA_OK(cudaFuncSetAttribute(csmri_fused<R, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_set |= 1ull << (dev & 63);
  }
  TFPNP_CUDA_OK(launch_ex(csmri_fused<R, CL>, dim3(CL * B), dim3(FUSED_WARPS * 32), smem, st, true, CL, x, z, u, d, y0p, maskp, mu));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

template <int R>
int launch_update(const float* x, float2* z, float2* u, float* d, float2* T, const float2* y0p,
                  const uint8_t* maskp, const float* mu, int B, cudaStream_t st) {
  constexpr int N = 32 * R;
  if constexpr (R == 2 || R == 4) {
    // Opt-in (TFPNP_CSMRI_FUSED=1).  Measured on B200 at 48 x 128^2: the update segment drops from 34 to 29 us per
    // iteration, but the whole step gets 1.9 % SLOWER (24.39 vs 23.93 ms): 96 clusters with 132 KB of shared memory each
    // and scattered 8-byte DSMEM stores hold up the start of the next denoiser call more than the two saved launches
    // give back.  Kept (tested against the three-kernel path) as the base for a bulk-DSMEM transpose.
    static const int fused = getenv("TFPNP_CSMRI_FUSED") ? atoi(getenv("TFPNP_CSMRI_FUSED")) : 0;
    if (fused == 2) return launch_fused<R, 2>(x, z, u, d, y0p, maskp, mu, B, st);
    if (fused == 4) return launch_fused<R, R == 4 ? 4 : 2>(x, z, u, d, y0p, maskp, mu, B, st);
  }
  const int row_blocks = B * N / ROWS_PER_CTA;
  // TFPNP_CSMRI_PDLMASK (bit 0 rows_fwd, 1 cols, 2 rows_inv) launches the kernels with programmatic stream serialisation;
  // each executes griddepcontrol.wait before it touches its predecessor's output.  Off by default: measured 64.4k vs
  // 65.1k image-iterations/s (fp16) and 29.3k vs 29.7k (fp16x3) with all three on -- the early-launched CTAs take SM
  // slots from the denoiser's last layer and the update is three round trips to L2 either way.
  static const bool pdl = !(getenv("TFPNP_PDL") && atoi(getenv("TFPNP_PDL")) == 0);
  static const int pm = getenv("TFPNP_CSMRI_PDLMASK") ? atoi(getenv("TFPNP_CSMRI_PDLMASK")) : 0;
  TFPNP_CUDA_OK(launch_ex(csmri_rows_fwd<R>, dim3(row_blocks), dim3(ROWS_PER_CTA * 32), 0, st, pdl && (pm & 1), 1, x, u, T));
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(launch_ex(csmri_cols<R>, dim3(N / COLS_PER_CTA, B), dim3(COLS_PER_CTA * 32), 0, st, pdl && (pm & 2), 1, T, y0p, maskp, mu));
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(launch_ex(csmri_rows_inv<R>, dim3(row_blocks), dim3(ROWS_PER_CTA * 32), 0, st, pdl && (pm & 4), 1, T, x, z, u, d));
  TFPNP_COUNT_LAUNCH();
  return 0;
}

}  // namespace

int csmri_prep(const float* y0, const uint8_t* mask, float2* y0p, uint8_t* maskp, int B, int N,
               cudaStream_t st) {
  TFPNP_CUDA_OK(fft_tables_init());   // twiddles: once per device, outside any graph capture
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "csmri: N must be 32/64/128/256, got %d", N);
  size_t n = (size_t)B * N * N;
  csmri_prep_kernel<<<(unsigned)(n / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(y0), mask, y0p,
                                                         maskp, N, N / 32);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int csmri_update(const float* x, float2* z, float2* u, float* d, float2* T, const float2* y0p,
                 const uint8_t* maskp, const float* mu, int B, int N, cudaStream_t st) {
  switch (N) {
    case 32: return launch_update<1>(x, z, u, d, T, y0p, maskp, mu, B, st);
    case 64: return launch_update<2>(x, z, u, d, T, y0p, maskp, mu, B, st);
    case 128: return launch_update<4>(x, z, u, d, T, y0p, maskp, mu, B, st);
    case 256: return launch_update<8>(x, z, u, d, T, y0p, maskp, mu, B, st);
  }
  set_error("csmri: unsupported size %d", N);
  return TFPNP_ERR_INVALID;
}

}  // namespace tfpnp
