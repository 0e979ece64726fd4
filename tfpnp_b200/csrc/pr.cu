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
#include <cstdlib>
#include "tasks.cuh"
#include "fft.cuh"
#include "fft256.cuh"
#include "sm100.cuh"

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


// ---- N = 256: half-warp transforms, 16 points per lane (fft256.cuh) -----------------------------------------------------------
// Same three passes; frequencies stay in natural order, so y0p is just |y0| transposed to [B][M][col][row].
constexpr int kHw = 16;   // half-warps (= transforms) per 256-thread CTA

__global__ void pr256_prep_kernel(const float* __restrict__ y0, float* __restrict__ y0p) {
  __shared__ float t[32][33];
  const size_t bm = blockIdx.z;
  const float* src = y0 + bm * 65536;
  float* dst = y0p + bm * 65536;
  const int x0 = blockIdx.x * 32, y0_ = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) t[i][threadIdx.x] = src[(size_t)(y0_ + i) * 256 + x0 + threadIdx.x];
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) dst[(size_t)(x0 + i) * 256 + y0_ + threadIdx.x] = t[threadIdx.x][i];
}

__global__ void __launch_bounds__(kHw * 16)
pr256_rows_fwd(const float2* __restrict__ z, const float2* __restrict__ mask, float2* __restrict__ T, int M) {
  __shared__ float2 s_tw[256];
  __shared__ float2 s_x[kHw][kF256Slots];
  s_tw[threadIdx.x] = g_fft_tw256[threadIdx.x];
  __syncthreads();
  const int hw = threadIdx.x >> 4, t = threadIdx.x & 15;
  const size_t row = (size_t)blockIdx.x * kHw + hw;   // over (b, j, r)
  const int r = row & 255;
  const size_t b = row / ((size_t)256 * M);
  const float2* zr = z + (b * 256 + r) * 256;
  const float2* mr = mask + row * 256;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = cmul(zr[t + 16 * j], __ldcs(mr + t + 16 * j));   // masks: streamed (evict-first), T stays in L2
  fft256_run<false, 1>(v, t, s_x[hw], s_tw);
  float2* tr = T + row * 256;
#pragma unroll
  for (int j = 0; j < 16; ++j) tr[t + 16 * j] = v[j];
}

// C columns of one (image, mask) spectrum per CTA, one half-warp per column.  The tile keeps element n of a column at row
// n + (n >> 4): exactly the exchange slots of fft256_run with stride = pitch, so the transforms run in place.
template <int C, int OCC>
__global__ void __launch_bounds__(C * 16, OCC * 256 / (C * 16))
pr256_cols(float2* __restrict__ T, const float* __restrict__ y0p) {
  constexpr int PITCH = C + 1;
  __shared__ float2 s_tw[256];
  __shared__ float2 tile[kF256Slots * PITCH];
  for (int i = threadIdx.x; i < 256; i += C * 16) s_tw[i] = g_fft_tw256[i];
  const size_t bm = blockIdx.y;
  const int c0 = blockIdx.x * C;
  float2* Tb = T + bm * 65536;
#pragma unroll
  for (int q = 0; q < 16; ++q) {                    // 256 * C elements / (16 C threads): 16 independent loads per thread
    const int i = threadIdx.x + q * C * 16;
    const int r = i / C, cc = i % C;
    tile[(r + (r >> 4)) * PITCH + cc] = Tb[(size_t)r * 256 + c0 + cc];
  }
  const int w = threadIdx.x >> 4, t = threadIdx.x & 15;
  const float* yc = y0p + (bm * 256 + c0 + w) * 256;
  float yv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) yv[j] = __ldcs(yc + t + 16 * j);   // in flight during the forward transform
  __syncthreads();
  float2* base = tile + w;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = base[(17 * j + t) * PITCH];
  fft256_run<false, PITCH>(v, t, base, s_tw);
  const float inv_n = 1.0f / 256.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 a = make_float2(v[j].x * inv_n, v[j].y * inv_n);     // Az
    float yh = sqrtf(a.x * a.x + a.y * a.y);                    // complex_abs, transforms.py:118
    float ratio = (yh - yv[j]) / yh;                            // meas_err / y_hat, solver.py:66-67
    v[j] = make_float2(ratio * a.x, ratio * a.y);
  }
  fft256_run<true, PITCH>(v, t, base, s_tw);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) base[(17 * j + t) * PITCH] = v[j];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const int i = threadIdx.x + q * C * 16;
    const int r = i / C, cc = i % C;
    Tb[(size_t)r * 256 + c0 + cc] = tile[(r + (r >> 4)) * PITCH + cc];
  }
}

template <int OCC>
__global__ void __launch_bounds__(kHw * 16, OCC)
pr256_rows_inv(const float2* __restrict__ T, const float2* __restrict__ mask, const float* __restrict__ x,
               float2* __restrict__ z, float2* __restrict__ u, float* __restrict__ d,
               const float* __restrict__ mu, const float* __restrict__ tau, int M) {
  __shared__ float2 s_tw[256];
  __shared__ float2 s_x[kHw][kF256Slots];
  s_tw[threadIdx.x] = g_fft_tw256[threadIdx.x];
  __syncthreads();
  // four rows per CTA, four half-warps per row: half-warp (ri, ml) inverts the spectra of masks ml, ml + 4, ... of its row,
  // the four partial sums meet in shared memory.  (One half-warp looping over the masks left the SM waiting on each load.)
  const int hw = threadIdx.x >> 4, t = threadIdx.x & 15;
  const int ri = hw >> 2, ml = hw & 3;
  const size_t row = (size_t)blockIdx.x * 4 + ri;     // over (b, r)
  const int r = row & 255;
  const size_t b = row >> 8;
  const float inv_n = 1.0f / 256.0f;
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = make_float2(0.f, 0.f);
  // the trip count is uniform over the CTA (fft256_run synchronises the whole warp and its two half-warps own different
  // masks): a half-warp whose mask index runs past M transforms zeros -- with 1 or 3 masks the two halves of a warp would
  // otherwise execute different numbers of __syncwarp and hang
  for (int m0 = 0; m0 < M; m0 += 4) {
    const int m = m0 + ml;
    const bool live = m < M;
    const size_t mrow = ((b * M + (live ? m : 0)) * 256 + r) * 256;
    float2 acc[16], mk[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = v[j];       // (first trip: zero)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = live ? T[mrow + t + 16 * j] : make_float2(0.f, 0.f);
#pragma unroll
    for (int j = 0; j < 16; ++j) mk[j] = live ? __ldcs(mask + mrow + t + 16 * j) : make_float2(0.f, 0.f);
    fft256_run<true, 1>(v, t, s_x[hw], s_tw);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float2 q = make_float2(v[j].x * inv_n, v[j].y * inv_n);
      v[j] = cadd(acc[j], cmulc(q, mk[j]));   // * conj(mask), transforms.py:319
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 16; ++j) s_x[hw][t + 16 * j] = v[j];
  __syncthreads();
  const int fr = threadIdx.x >> 6;                      // finalise: 64 threads per row, coalesced
  const size_t frow = (size_t)blockIdx.x * 4 + fr;
  const size_t fb = frow >> 8;
  const float mu_b = mu[fb], tau_b = tau[fb], inv_m = 1.0f / (float)M;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int px = (threadIdx.x & 63) + 64 * q;
    float2 g = s_x[fr * 4][px];
#pragma unroll
    for (int k = 1; k < 4; ++k) g = cadd(g, s_x[fr * 4 + k][px]);
    g = make_float2(g.x * inv_m, g.y * inv_m);                         // .mean(1), transforms.py:320
    const size_t i = frow * 256 + px;
    float2 zz = z[i], uu = __ldcs(u + i);
    float xx = __ldcs(x + i);
    zz.x = zz.x - tau_b * (g.x + mu_b * (zz.x - (xx + uu.x)));        // solver.py:69
    zz.y = zz.y - tau_b * (g.y + mu_b * (zz.y - uu.y));
    uu.x = uu.x + xx - zz.x;                                           // solver.py:72
    uu.y = uu.y - zz.y;
    z[i] = zz;
    __stcs(u + i, uu);
    d[i] = zz.x - uu.x;
  }
}


// The column pass with its tile staged by TMA: ONE cp.async.bulk.tensor load brings the 256 x 16 float2 tile of an (image, mask)
// spectrum into shared memory (128-byte rows, SWIZZLE_128B: element (r, w) sits in 16-byte chunk (w/2) ^ (r & 7) of row r, so
// a column walk by 16 lanes touches all 32 banks twice), the half-warps transform their columns through separate exchange
// slots, write the result back into the tile and ONE bulk tensor store returns it.  Replaces 2 x 16 (LDG/STG + STS/LDS +
// index arithmetic) per thread of pr256_cols.
// 2-D TMA tile moves (this file's only use of them): global -> shared completing on an mbarrier, shared -> global as a bulk group
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          sm100::smem_u32(smem_dst)),
      "l"(map), "r"(sm100::smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(sm100::smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

constexpr int kColsTmaTile = 256 * 128;                                             // bytes
constexpr int kColsTmaSmem = 1024 + kColsTmaTile + 16 * kF256Slots * 8 + 256 * 8 + 16;   // align slack, tile, slots, twiddles, barrier

__global__ void __launch_bounds__(256, 2)
pr256_cols_tma(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ y0p) {
  extern __shared__ uint8_t cols_raw[];
  uint8_t* tile = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(cols_raw) + 1023) & ~(uintptr_t)1023);
  float2* s_x = reinterpret_cast<float2*>(tile + kColsTmaTile);
  float2* s_tw = s_x + 16 * kF256Slots;
  uint64_t* bar = reinterpret_cast<uint64_t*>(s_tw + 256);
  const int tid = threadIdx.x;
  const size_t bm = blockIdx.y;
  const int c0 = blockIdx.x * 16;
  if (tid == 0) {
    sm100::mbar_init(bar, 1);
    sm100::fence_barrier_init();
  }
  s_tw[tid] = g_fft_tw256[tid];
  __syncthreads();
  if (tid == 0) {
    sm100::mbar_arrive_expect_tx(bar, kColsTmaTile);
    tma_load_2d(tile, &tmap, bar, 2 * c0, (int)(bm * 256));   // coordinates in fp32 elements (2 per float2), rows
  }
  const int w = tid >> 4, t = tid & 15;
  const float* yc = y0p + (bm * 256 + c0 + w) * 256;
  float yv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) yv[j] = __ldcs(yc + t + 16 * j);   // in flight during the tile load and the forward transform
  sm100::mbar_wait(bar, 0);
  // (r & 7) == (t & 7) for r = t + 16 j: the chunk of this lane's element is the same in every row it touches
  uint8_t* lane_base = tile + t * 128 + ((((w >> 1) ^ (t & 7)) << 4) | ((w & 1) << 3));
  float2 v[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = *reinterpret_cast<const float2*>(lane_base + j * (16 * 128));
  fft256_run<false, 1>(v, t, s_x + w * kF256Slots, s_tw);
  const float inv_n = 1.0f / 256.0f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 a = make_float2(v[j].x * inv_n, v[j].y * inv_n);     // Az
    float yh = sqrtf(a.x * a.x + a.y * a.y);                    // complex_abs, transforms.py:118
    float ratio = (yh - yv[j]) / yh;                            // meas_err / y_hat, solver.py:66-67
    v[j] = make_float2(ratio * a.x, ratio * a.y);
  }
  fft256_run<true, 1>(v, t, s_x + w * kF256Slots, s_tw);
#pragma unroll
  for (int j = 0; j < 16; ++j) *reinterpret_cast<float2*>(lane_base + j * (16 * 128)) = v[j];
  sm100::fence_proxy_async();        // generic-proxy writes to the tile -> visible to the bulk store
  __syncthreads();
  if (tid == 0) {
    tma_store_2d(&tmap, tile, 2 * c0, (int)(bm * 256));
    tma_store_commit_and_wait();
  }
}

// 2-D fp32 view of the spectra workspace T: [B*M*256 rows][512 floats], box 32 floats x 256 rows, 128-byte swizzle
int pr256_encode_map(CUtensorMap* m, float2* T, int BM) {
  typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    TFPNP_CHECK(e == cudaSuccess && q == cudaDriverEntryPointSuccess && p, "cuTensorMapEncodeTiled entry point unavailable");
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t dims[2] = {512, (cuuint64_t)BM * 256};
  const cuuint64_t strides[1] = {2048};
  const cuuint32_t box[2] = {32, 256}, estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, T, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  TFPNP_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (pr spectra, %d rows)", (int)r, BM * 256);
  return 0;
}

int pr256_cols_use_tma() {
  static const int v = getenv("TFPNP_PR_COLS_TMA") ? atoi(getenv("TFPNP_PR_COLS_TMA")) : 1;
  return v;
}

// TFPNP_PR_FFT16=0 keeps the warp-wide transform at N = 256 (A/B switch; read once)
bool pr_use_fft16() {
  static const int v = getenv("TFPNP_PR_FFT16") ? atoi(getenv("TFPNP_PR_FFT16")) : 1;
  return v != 0;
}
int pr256_cols_width() {
  static const int v = getenv("TFPNP_PR_COLS") ? atoi(getenv("TFPNP_PR_COLS")) : 16;
  return v;
}

int pr256_occ() {   // tens: CTAs of 256 threads per SM the column kernel is compiled for, units: same for rows_inv
  static const int v = getenv("TFPNP_PR_OCC") ? atoi(getenv("TFPNP_PR_OCC")) : 22;
  return v;
}
int pr256_chunk() {
  static const int v = getenv("TFPNP_PR_CHUNK") ? atoi(getenv("TFPNP_PR_CHUNK")) : 0;
  return v;
}

// TFPNP_PR_CHUNK=n runs the three passes over chunks of n images so that one chunk's spectra stay in L2 (experiment, off:
// 36 images in chunks of 9 / 6 / 3 measured 254 / 272 / 495 us per iteration against 208 un-chunked -- the passes are bound by
// per-launch latency, not by HBM bytes).
int launch_update256(const float* x, float2* z, float2* u, float* d, float2* T, const float* y0p,
                     const float2* mask, const float* mu, const float* tau, int B, int M, cudaStream_t st) {
  const int chunk = pr256_chunk() > 0 ? pr256_chunk() : B;
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int nb = B - b0 < chunk ? B - b0 : chunk;
    const size_t oi = (size_t)b0 * 65536, om = oi * M;
    pr256_rows_fwd<<<nb * M * 256 / kHw, kHw * 16, 0, st>>>(z + oi, mask + om, T + om, M);
    TFPNP_COUNT_LAUNCH();
    const int occ = pr256_occ();
    const dim3 g8(256 / 8, nb * M), g16(256 / 16, nb * M);
    if (pr256_cols_use_tma() && chunk >= B) {
      static unsigned long long attr_set = 0;
      int dev = 0;
      cudaGetDevice(&dev);
      if (!((attr_set >> (dev & 63)) & 1ull)) {
        TFPNP_CUDA_OK(cudaFuncSetAttribute(pr256_cols_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, kColsTmaSmem));
        attr_set |= 1ull << (dev & 63);
      }
      CUtensorMap tmap;
      TFPNP_TRY(pr256_encode_map(&tmap, T, B * M));
      pr256_cols_tma<<<g16, 256, kColsTmaSmem, st>>>(tmap, y0p);
    } else if (pr256_cols_width() == 8) {
      if (occ / 10 == 3) pr256_cols<8, 3><<<g8, 8 * 16, 0, st>>>(T + om, y0p + om);
      else pr256_cols<8, 2><<<g8, 8 * 16, 0, st>>>(T + om, y0p + om);
    } else {
      if (occ / 10 == 3) pr256_cols<16, 3><<<g16, 16 * 16, 0, st>>>(T + om, y0p + om);
      else pr256_cols<16, 2><<<g16, 16 * 16, 0, st>>>(T + om, y0p + om);
    }
    TFPNP_COUNT_LAUNCH();
    if (occ % 10 == 3)
      pr256_rows_inv<3><<<nb * 256 / 4, kHw * 16, 0, st>>>(T + om, mask + om, x + oi, z + oi, u + oi, d + oi, mu + b0, tau + b0, M);
    else
      pr256_rows_inv<2><<<nb * 256 / 4, kHw * 16, 0, st>>>(T + om, mask + om, x + oi, z + oi, u + oi, d + oi, mu + b0, tau + b0, M);
    TFPNP_COUNT_LAUNCH();
  }
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
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
  if (N == 256 && pr_use_fft16()) pr256_prep_kernel<<<dim3(8, 8, B * M), dim3(32, 8), 0, st>>>(y0, y0p);
  else pr_prep_kernel<<<(unsigned)(n / 256), 256, 0, st>>>(y0, y0p, N, N / 32);
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
    case 256:
      if (pr_use_fft16()) return launch_update256(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
      return launch_update<8>(x, z, u, d, T, y0p, mask, mu, tau, B, M, st);
  }
  set_error("pr: unsupported size %d", N);
  return TFPNP_ERR_INVALID;
}

}  // namespace tfpnp

// ---- reverse mode: sequence and element bodies in grad_elem.cuh; FFTs through tfpnp_fft2 (fft_ops.cu) --------
namespace tfpnp {
namespace {

__global__ void prg_slot_copy(const float2* __restrict__ state, float2* __restrict__ buf, int k, int HW, size_t n, int to_state) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2* s = const_cast<float2*>(state) + ((i / HW) * 3 + k) * HW + i % HW;
  if (to_state) *s = buf[i]; else buf[i] = *s;
}
__global__ void prg_pre(const float2* __restrict__ GZ, const float2* __restrict__ GU, const float2* __restrict__ st_i,
                        float2* __restrict__ GZT, float2* __restrict__ Z, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_pre_elem(i, GZ, GU, st_i, GZT, Z, HW);
}
__global__ void prg_mul(const float2* __restrict__ img, const float2* __restrict__ mask, float2* __restrict__ out, int M, int HW,
                        size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_mul_elem(i, img, mask, out, M, HW);
}
__global__ void prg_h(float2* __restrict__ W, float2* __restrict__ Bc, const float* __restrict__ y0, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_h_elem(i, W, Bc, y0);
}
__global__ void prg_acc(const float2* __restrict__ E, const float2* __restrict__ mask, float2* __restrict__ out, int M, int HW,
                        size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_acc_elem(i, E, mask, out, M, HW);
}
__global__ void prg_mid(const float2* __restrict__ st_i, const float2* __restrict__ st_n, const float2* __restrict__ GZT,
                        const float2* __restrict__ JC, const float2* __restrict__ Gz, const float* __restrict__ mu,
                        const float* __restrict__ tau, const float2* __restrict__ GX, float2* __restrict__ GZ,
                        float2* __restrict__ GU, float* __restrict__ gxt, float* __restrict__ v, float* __restrict__ t_tau,
                        float* __restrict__ t_mu, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_mid_elem(i, st_i, st_n, GZT, JC, Gz, mu, tau, GX, GZ, GU, gxt, v, t_tau, t_mu, HW);
}
__global__ void prg_post(const float* __restrict__ gv, float2* __restrict__ GX, float2* __restrict__ GZ, float2* __restrict__ GU,
                         size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::pr_post_elem(i, gv, GX, GZ, GU);
}
__global__ void __launch_bounds__(256)
prg_image_sum(const float* __restrict__ term, float* __restrict__ out, int64_t stride, int HW) {
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
__global__ void prg_gather_params(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                                  int64_t rs, int64_t cs, float* __restrict__ P, int B, int iters) {
  const int n = B * iters;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = t / B, b = t % B;
    const int64_t src = b * rs + i * cs;
    P[t] = p0[src]; P[n + t] = p1[src]; P[2 * n + t] = p2[src];
  }
}

struct PrGradOps {
  Denoiser* den; const float2* mask; const float* y0; float2* fft_ws; int B, M, N; cudaStream_t st;
  static constexpr int T = 256;
  int HW() const { return N * N; }
  size_t n() const { return (size_t)B * HW(); }
  size_t nm() const { return n() * M; }
  unsigned nb(size_t k) const { return (unsigned)((k + T - 1) / T); }
  int slot_get(const float2* state, float2* buf, int k) {
    prg_slot_copy<<<nb(n()), T, 0, st>>>(state, buf, k, HW(), n(), 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int slot_put(float2* state, float2* buf, int k) {
    prg_slot_copy<<<nb(n()), T, 0, st>>>(state, buf, k, HW(), n(), 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int pre(const float2* gz, const float2* gu, const float2* st_i, float2* gzt, float2* z) {
    prg_pre<<<nb(n()), T, 0, st>>>(gz, gu, st_i, gzt, z, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int mul(const float2* img, float2* out) {
    prg_mul<<<nb(nm()), T, 0, st>>>(img, mask, out, M, HW(), nm());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int fft(const float2* in, float2* out, bool inverse) {
    TFPNP_COUNT_LAUNCH(); TFPNP_COUNT_LAUNCH();
    return tfpnp_fft2(reinterpret_cast<const float*>(in), reinterpret_cast<float*>(out), reinterpret_cast<float*>(fft_ws), B * M, N,
                      inverse ? 1 : 0, 0, st);
  }
  int h(float2* W, float2* Bc) {
    prg_h<<<nb(nm()), T, 0, st>>>(W, Bc, y0, nm());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int acc(const float2* E, float2* out) {
    prg_acc<<<nb(n()), T, 0, st>>>(E, mask, out, M, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int mid(const float2* st_i, const float2* st_n, const float2* gzt, const float2* jc, const float2* gzv, const float* mu_i,
          const float* tau_i, const float2* gx, float2* gz, float2* gu, float* gxt, float* v, float* t_tau, float* t_mu) {
    prg_mid<<<nb(n()), T, 0, st>>>(st_i, st_n, gzt, jc, gzv, mu_i, tau_i, gx, gz, gu, gxt, v, t_tau, t_mu, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    prg_image_sum<<<B, 256, 0, st>>>(term, out, stride, HW());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return den->vjp(v, sg_i, 1, gxt, gv, gsig, stride, B, N, N, st);
  }
  int post(const float* gv, float2* gx, float2* gz, float2* gu) {
    prg_post<<<nb(n()), T, 0, st>>>(gv, gx, gz, gu, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
};

}  // namespace
}  // namespace tfpnp

extern "C" int tfpnp_pr_iadmm_backward(void* denoiser, const float* states, const float* y0, const float* mask, int n_masks,
                                       const float* sigma_d, const float* mu, const float* tau, int64_t row_stride,
                                       int64_t col_stride, int B, int N, int iters, const float* grad_out, float* grad_sigma_d,
                                       float* grad_mu, float* grad_tau, float* grad_state_in, void* stream) {
  using namespace tfpnp;
  TFPNP_CHECK(denoiser && states && y0 && mask && sigma_d && mu && tau && grad_out && grad_sigma_d && grad_mu && grad_tau &&
                  B > 0 && iters > 0 && n_masks > 0, "bad argument");
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "FFT tasks support N in {32,64,128,256}, got %d", N);
  g_launch_count = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n = (size_t)B * N * N, nm = n * n_masks;
  PoolBuf c1[7], cm[5], f1[5], P;
  auto body = [&]() -> int {
    for (PoolBuf& b : c1) TFPNP_TRY(b.alloc(n * sizeof(float2), st));
    for (PoolBuf& b : cm) TFPNP_TRY(b.alloc(nm * sizeof(float2), st));
    for (PoolBuf& b : f1) TFPNP_TRY(b.alloc(n * sizeof(float), st));
    TFPNP_TRY(P.alloc((size_t)B * iters * 3 * sizeof(float), st));
    prg_gather_params<<<cdiv(B * iters, 256), 256, 0, st>>>(sigma_d, mu, tau, row_stride, col_stride, P.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    PrGradOps ops{static_cast<Denoiser*>(denoiser), reinterpret_cast<const float2*>(mask), y0, cm[4].as<float2>(), B, n_masks, N, st};
    grad_elem::PrGradBufs w{c1[0].as<float2>(), c1[1].as<float2>(), c1[2].as<float2>(), c1[3].as<float2>(), c1[4].as<float2>(),
                            c1[5].as<float2>(), c1[6].as<float2>(), cm[0].as<float2>(), cm[1].as<float2>(), cm[2].as<float2>(),
                            cm[3].as<float2>(), f1[0].as<float>(), f1[1].as<float>(), f1[2].as<float>(), f1[3].as<float>(),
                            f1[4].as<float>()};
    TFPNP_TRY(grad_elem::pr_backward_sequence(ops, reinterpret_cast<const float2*>(states), P.as<float>(), B, N * N, iters,
                                              reinterpret_cast<const float2*>(grad_out), grad_sigma_d, grad_mu, grad_tau,
                                              reinterpret_cast<float2*>(grad_state_in), w));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  };
  const int rc = body();
  for (PoolBuf& b : c1) b.release();
  for (PoolBuf& b : cm) b.release();
  for (PoolBuf& b : f1) b.release();
  P.release();
  return rc;
}
