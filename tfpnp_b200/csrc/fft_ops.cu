// Stand-alone 2-D transforms of tfpnp/utils/transforms.py as device operators:
//   fft2 / ifft2 (:68-103)  centred, ortho-normalised:  fftshift(FFT2_ortho(ifftshift(x)))  over dims (-3,-2)
//   the un-centred ortho pair used by cdp_forward / cdp_backward (:282-320)
// The solvers never call these (their FFTs are fused with the data-fidelity steps, csmri.cu / pr.cu); they serve the
// measurement synthesis on the GPU (y0 = fft2(gt) + noise, |cdp_forward(gt)|, ...) and pin rows T2/T3 to the reference's
// golden vectors on their own.  Built from the same warp-register FFT (fft.cuh): a row pass (warp per row) and a column
// pass (8 columns per CTA through a padded shared-memory tile).  The forward transform leaves its output in the
// bit-reversed "P order" of fft.cuh, the inverse consumes it; the natural-order <-> P-order permutation and the
// centring rolls (for even N: both a roll by N/2) are folded into the index arithmetic of the loads / stores.
#include "tasks.cuh"
#include "fft.cuh"

namespace tfpnp {
namespace {

constexpr int FO_ROWS_PER_CTA = 8;
constexpr int FO_COLS_PER_CTA = 8;

// forward rows: T[b][r][p] = rowFFT(x'[r][:])[freq(p)],  x'[r][c] = in[b][(r+sh)%N][(c+sh)%N]
// inverse rows: T[b][r][n] = N * rowIFFT(X'[r][:])[n],   X'[r][k] = in[b][(r+sh)%N][(k+sh)%N]  (presented in P order)
template <int R, bool INV>
__global__ void __launch_bounds__(FO_ROWS_PER_CTA * 32)
fft2_rows(const float2* __restrict__ in, float2* __restrict__ T, int sh) {
  constexpr int N = 32 * R;
  WarpFFT<R> f;
  f.init();
  const size_t row = (size_t)blockIdx.x * FO_ROWS_PER_CTA + (threadIdx.x >> 5);     // over (b, r)
  const int r = (int)(row % N);
  const size_t b = row / N;
  const float2* src = in + (b * N + (size_t)((r + sh) % N)) * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int pos = 32 * j + f.lane;
    const int c = INV ? fft_pos_to_freq(pos, R) : pos;
    v[j] = src[(c + sh) % N];
  }
  if (INV) f.inverse(v); else f.forward(v);
  float2* dst = T + row * N;
#pragma unroll
  for (int j = 0; j < R; ++j) dst[32 * j + f.lane] = v[j];
}

// forward cols: out[b][(kr+sh)%N][(kc+sh)%N] = colFFT(T[b][:][p])[kr] / N,  kc = freq(p)
// inverse cols: out[b][(m+sh)%N][(c+sh)%N]   = colIFFT(T[b][:][c])[m] / N   (T rows are frequencies in natural order)
template <int R, bool INV>
__global__ void __launch_bounds__(FO_COLS_PER_CTA * 32)
fft2_cols(const float2* __restrict__ T, float2* __restrict__ out, int sh) {
  constexpr int N = 32 * R;
  constexpr int PITCH = FO_COLS_PER_CTA + 1;
  __shared__ float2 tile[N * PITCH];
  const size_t b = blockIdx.y;
  const int c0 = blockIdx.x * FO_COLS_PER_CTA;
  const float2* Tb = T + b * N * N;
  for (int i = threadIdx.x; i < N * FO_COLS_PER_CTA; i += FO_COLS_PER_CTA * 32) {
    const int r = i / FO_COLS_PER_CTA, cc = i % FO_COLS_PER_CTA;
    tile[r * PITCH + cc] = Tb[(size_t)r * N + c0 + cc];
  }
  __syncthreads();
  WarpFFT<R> f;
  f.init();
  const int w = threadIdx.x >> 5;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int pos = 32 * j + f.lane;
    v[j] = tile[(INV ? fft_pos_to_freq(pos, R) : pos) * PITCH + w];
  }
  if (INV) f.inverse(v); else f.forward(v);
  const float inv_n = 1.0f / (float)N;          // ortho 2-D scale (1/sqrt(N) per axis), exact power of two
  float2* ob = out + b * N * N;
  const int col = INV ? c0 + w : fft_pos_to_freq(c0 + w, R);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const int pos = 32 * j + f.lane;
    const int rowi = INV ? pos : fft_pos_to_freq(pos, R);
    ob[(size_t)((rowi + sh) % N) * N + (col + sh) % N] = make_float2(v[j].x * inv_n, v[j].y * inv_n);
  }
}

template <int R>
int launch_fft2(const float2* in, float2* out, float2* T, int n_imgs, bool inverse, int sh, cudaStream_t st) {
  constexpr int N = 32 * R;
  const int row_blocks = n_imgs * N / FO_ROWS_PER_CTA;
  if (inverse) {
    fft2_rows<R, true><<<row_blocks, FO_ROWS_PER_CTA * 32, 0, st>>>(in, T, sh);
    fft2_cols<R, true><<<dim3(N / FO_COLS_PER_CTA, n_imgs), FO_COLS_PER_CTA * 32, 0, st>>>(T, out, sh);
  } else {
    fft2_rows<R, false><<<row_blocks, FO_ROWS_PER_CTA * 32, 0, st>>>(in, T, sh);
    fft2_cols<R, false><<<dim3(N / FO_COLS_PER_CTA, n_imgs), FO_COLS_PER_CTA * 32, 0, st>>>(T, out, sh);
  }
  TFPNP_COUNT_LAUNCH();
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace tfpnp

using namespace tfpnp;

extern "C" int tfpnp_fft2(const float* in, float* out, float* workspace, int n_imgs, int N, int inverse, int centered,
                          void* stream) {
  TFPNP_CHECK(in && out && workspace && n_imgs > 0, "tfpnp_fft2: bad argument");
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "tfpnp_fft2: N must be 32/64/128/256, got %d", N);
  TFPNP_CHECK(in != out && workspace != in && workspace != out, "tfpnp_fft2: in / out / workspace must not alias");
  TFPNP_CUDA_OK(fft_tables_init());
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float2* i2 = reinterpret_cast<const float2*>(in);
  float2* o2 = reinterpret_cast<float2*>(out);
  float2* t2 = reinterpret_cast<float2*>(workspace);
  const int sh = centered ? N / 2 : 0;
  switch (N) {
    case 32: return launch_fft2<1>(i2, o2, t2, n_imgs, inverse != 0, sh, st);
    case 64: return launch_fft2<2>(i2, o2, t2, n_imgs, inverse != 0, sh, st);
    case 128: return launch_fft2<4>(i2, o2, t2, n_imgs, inverse != 0, sh, st);
    default: return launch_fft2<8>(i2, o2, t2, n_imgs, inverse != 0, sh, st);
  }
}
