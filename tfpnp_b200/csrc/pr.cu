// Phase retrieval (coded diffraction patterns): inexact-ADMM z-step fused with the dual
// update (tasks/pr/solver.py:57-72, tfpnp/utils/transforms.py:106-118,260-320):
//     Az_j = FFT2_ortho(z * mask_j)                     (un-centred, j = 1..M)
//     g    = mean_j( IFFT2_ortho((|Az_j| - y0_j)/|Az_j| * Az_j) * conj(mask_j) )
//     z   -= tau (g + mu (z - (x + u)));   u += x - z;   d = Re(z - u)
//
// Three launches per iteration (the reference: repeat/stack/complex_mul copies + 2 cuFFT
// calls + ~20 elementwise launches).  The M masked spectra live in the workspace T
// (L2-resident at the BASELINE shapes); each pass is a coalesced float2 stream:
//   rows_fwd : warp per (image, mask, row)   z*mask_j -> row FFT -> T
//   cols     : 16 columns per CTA            col FFT -> magnitude projection -> inverse col FFT
//   rows_inv : warp per (image, row)         M inverse row FFTs -> * conj(mask_j) -> mean
//                                            -> z, u, d update
#include "tasks.cuh"
#include "fft.cuh"

namespace tfpnp {
namespace {

constexpr int ROWS_PER_CTA = 8;
constexpr int COLS_PER_CTA = 8;   // columns (= warps) per CTA in the column kernel: more, smaller CTAs per SM overlap the load / FFT / store phases

__global__ void pr_prep_kernel(const float* __restrict__ y0, float* __restrict__ y0p, int N, int R) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // (bm, c, r), r fastest
  int r = i % N, c = (i / N) % N;
  size_t bm = i / ((size_t)N * N);
  int kx = fft_pos_to_freq(c, R), ky = fft_pos_to_freq(r, R);
  y0p[i] = y0[(bm * N + ky) * N + kx];
}

template <int R>
__global__ void __launch_bounds__(ROWS_PER_CTA * 32)
pr_rows_fwd(const float2* __restrict__ z, const float2* __restrict__ mask, float2* __restrict__ T, int M) {
  constexpr int N = 32 * R;
  WarpFFT<R> f;
  f.init();
  size_t row = (size_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);   // over (b, j, r)
  int r = row % N;
  size_t b = row / ((size_t)N * M);
  const float2* zr = z + (b * N + r) * N;
  const float2* mr = mask + row * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = cmul(zr[32 * j + f.lane], mr[32 * j + f.lane]);
  f.forward(v);
  float2* tr = T + row * N;
#pragma unroll
  for (int j = 0; j < R; ++j) tr[32 * j + f.lane] = v[j];
}

template <int R>
__global__ void __launch_bounds__(COLS_PER_CTA * 32)
pr_cols(float2* __restrict__ T, const float* __restrict__ y0p) {
  constexpr int N = 32 * R;
  constexpr int PITCH = COLS_PER_CTA + 1;
  __shared__ float2 tile[N * PITCH];
  const size_t bm = blockIdx.y;
  const int c0 = blockIdx.x * COLS_PER_CTA;
  float2* Tb = T + bm * N * N;
  for (int i = threadIdx.x; i < N * COLS_PER_CTA; i += COLS_PER_CTA * 32) {
    int r = i / COLS_PER_CTA, cc = i % COLS_PER_CTA;
    tile[r * PITCH + cc] = Tb[(size_t)r * N + c0 + cc];
  }
  __syncthreads();
  WarpFFT<R> f;
  f.init();
  const int w = threadIdx.x >> 5;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = tile[(32 * j + f.lane) * PITCH + w];
  f.forward(v);
  const float inv_n = 1.0f / (float)N;
  const size_t col = (bm * N + c0 + w) * N;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 a = make_float2(v[j].x * inv_n, v[j].y * inv_n);     // Az
    float yh = sqrtf(a.x * a.x + a.y * a.y);                    // complex_abs, transforms.py:118
    float ratio = (yh - y0p[col + 32 * j + f.lane]) / yh;       // meas_err / y_hat, solver.py:66-67
    v[j] = make_float2(ratio * a.x, ratio * a.y);
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
pr_rows_inv(const float2* __restrict__ T, const float2* __restrict__ mask, const float* __restrict__ x,
            float2* __restrict__ z, float2* __restrict__ u, float* __restrict__ d,
            const float* __restrict__ mu, const float* __restrict__ tau, int M) {
  constexpr int N = 32 * R;
  WarpFFT<R> f;
  f.init();
  size_t row = (size_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);   // over (b, r)
  int r = row % N;
  size_t b = row / N;
  float2 acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = make_float2(0.f, 0.f);
  const float inv_n = 1.0f / (float)N;
  for (int m = 0; m < M; ++m) {
    size_t mrow = ((b * M + m) * N + r) * N;
    float2 v[R];
#pragma unroll
    for (int j = 0; j < R; ++j) v[j] = T[mrow + 32 * j + f.lane];
    f.inverse(v);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      float2 t = make_float2(v[j].x * inv_n, v[j].y * inv_n);
      acc[j] = cadd(acc[j], cmulc(t, mask[mrow + 32 * j + f.lane]));   // * conj(mask), transforms.py:319
    }
  }
  const float mu_b = mu[b], tau_b = tau[b], inv_m = 1.0f / (float)M;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    size_t i = row * N + 32 * j + f.lane;
    float2 g = make_float2(acc[j].x * inv_m, acc[j].y * inv_m);        // .mean(1), transforms.py:320
    float2 zz = z[i], uu = u[i];
    float xx = x[i];
    // z = z - tau (g + mu (z - (x + u)))   (solver.py:69)
    zz.x = zz.x - tau_b * (g.x + mu_b * (zz.x - (xx + uu.x)));
    zz.y = zz.y - tau_b * (g.y + mu_b * (zz.y - uu.y));
    uu.x = uu.x + xx - zz.x;                                           // solver.py:72
    uu.y = uu.y - zz.y;
    z[i] = zz;
    u[i] = uu;
    d[i] = zz.x - uu.x;
  }
}

template <int R>
int launch_update(const float* x, float2* z, float2* u, float* d, float2* T, const float* y0p,
                  const float2* mask, const float* mu, const float* tau, int B, int M, cudaStream_t st) {
  constexpr int N = 32 * R;
  pr_rows_fwd<R><<<B * M * N / ROWS_PER_CTA, ROWS_PER_CTA * 32, 0, st>>>(z, mask, T, M);
  TFPNP_COUNT_LAUNCH();
  pr_cols<R><<<dim3(N / COLS_PER_CTA, B * M), COLS_PER_CTA * 32, 0, st>>>(T, y0p);
  TFPNP_COUNT_LAUNCH();
  pr_rows_inv<R><<<B * N / ROWS_PER_CTA, ROWS_PER_CTA * 32, 0, st>>>(T, mask, x, z, u, d, mu, tau, M);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int pr_prep(const float* y0, float* y0p, int B, int M, int N, cudaStream_t st) {
  TFPNP_CUDA_OK(fft_tables_init());   // twiddles: once per device, outside any graph capture
  size_t n = (size_t)B * M * N * N;
  pr_prep_kernel<<<(unsigned)(n / 256), 256, 0, st>>>(y0, y0p, N, N / 32);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int pr_update(const float* x, float2* z, float2* u, float* d, float2* T, const float* y0p,
              const float2* mask, const float* mu, const float* tau, int B, int M, int N, cudaStream_t st) {
  switch (N) {
    case 32: return launch_update<1>(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
    case 64: return launch_update<2>(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
    case 128: return launch_update<4>(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
    case 256: return launch_update<8>(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
  }
  set_error("pr: unsupported size %d", N);
  return TFPNP_ERR_INVALID;
}

}  // namespace tfpnp
