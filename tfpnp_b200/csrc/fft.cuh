// Warp-cooperative 1-D complex FFT of length N = 32*R (R in {1,2,4,8}) held in
// registers: lane l keeps v[j] = x[32*j + l].
//
//   forward : natural order in  -> "P order" out, v[k1] on lane l = X[k1 + R*bitrev5(l)]
//   inverse : "P order" in      -> natural order out (unscaled: N * ifft)
//
// Decomposition n = 32 j + l, k = k1 + R k2:
//   X[k1 + R k2] = sum_l W_N^{l k1} ( sum_j x[32 j + l] W_R^{j k1} ) W_32^{l k2}
// i.e. an R-point DFT in registers, a twiddle, and a 32-point radix-2 DIF across
// lanes with __shfl_xor (output bit-reversed over lanes).  The inverse undoes the
// stages in reverse (DIT) so no bit-reversal pass is ever needed: frequency-domain
// work is pointwise and its operands are pre-permuted once per solver call.
//
// Replaces torch.fft / torch.ifft (cuFFT) + the four narrow/cat rolls of
// tfpnp/utils/transforms.py:68-103,215-257.
#pragma once
#include <cuda_runtime.h>
#include <cmath>

namespace tfpnp {

__host__ __device__ inline int bitrev5(int l) {
  return ((l & 1) << 4) | ((l & 2) << 2) | (l & 4) | ((l & 8) >> 2) | ((l & 16) >> 4);
}
// frequency index stored at position c (register c/32, lane c%32) after a forward pass
__host__ __device__ inline int fft_pos_to_freq(int c, int R) { return (c >> 5) + R * bitrev5(c & 31); }

__host__ __device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__host__ __device__ __forceinline__ float2 cmulc(float2 a, float2 b) {  // a * conj(b)
  return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__host__ __device__ __forceinline__ float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

template <bool INV>
__host__ __device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
  float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi<INV>(csub(a1, a3));
  a0 = cadd(t0, t2);
  a2 = csub(t0, t2);
  a1 = cadd(t1, t3);
  a3 = csub(t1, t3);
}

template <int R, bool INV>
__device__ __forceinline__ void dft_regs(float2 (&v)[R]) {
  if constexpr (R == 2) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  } else if constexpr (R == 4) {
    dft4<INV>(v[0], v[1], v[2], v[3]);
  } else if constexpr (R == 8) {
    float2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
    float2 o0 = v[1], o1 = v[3], o2 = v[5], o3 = v[7];
    dft4<INV>(e0, e1, e2, e3);
    dft4<INV>(o0, o1, o2, o3);
    const float h = 0.70710678118654752440f;
    // W8^1 = h(1 -+ i), W8^2 = -+i, W8^3 = h(-1 -+ i)   (upper sign: forward)
    float2 w1 = INV ? make_float2(h * (o1.x - o1.y), h * (o1.x + o1.y))
                    : make_float2(h * (o1.x + o1.y), h * (o1.y - o1.x));
    float2 w2 = mul_mi<INV>(o2);
    float2 w3 = INV ? make_float2(-h * (o3.x + o3.y), h * (o3.x - o3.y))
                    : make_float2(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, w1); v[5] = csub(e1, w1);
    v[2] = cadd(e2, w2); v[6] = csub(e2, w2);
    v[3] = cadd(e3, w3); v[7] = csub(e3, w3);
  }
}

// Twiddle table, filled once per device from the host in double precision (fft_tables_init): a warp that does ONE
// transform spent as many instructions on its 5 + R sincospif calls as on the FFT itself.
//   g_fft_tw[log2 R][i][lane], i < 5 : W_{2s}^{lane & (s-1)}, s = 16 >> i ;   i = 5 + k : W_{32R}^{lane * k}
// NOTE: the table is `static` = one copy per translation unit (no relocatable device code in this build): every entry
// point of a .cu file that launches WarpFFT kernels must call fft_tables_init() itself (csmri_prep, pr_prep,
// tfpnp_fft2, tfpnp_csmri_variant_forward do).
constexpr int kFftTwRows = 13;
static __device__ float2 g_fft_tw[4][kFftTwRows][32];
static __device__ float2 g_fft_tw256[256];   // W_256^e = exp(-2 pi i e / 256): lane twiddles of the 16x16 transform (fft256.cuh)

static inline cudaError_t fft_tables_init() {
  static unsigned long long done_mask = 0;   // one bit per device (the table is per translation unit and device)
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if ((done_mask >> (dev & 63)) & 1ull) return cudaSuccess;
  static float2 h[4][kFftTwRows][32];
  const double pi = 3.14159265358979323846;
  for (int lr = 0; lr < 4; ++lr) {
    const int R = 1 << lr;
    for (int lane = 0; lane < 32; ++lane) {
      for (int i = 0; i < 5; ++i) {
        const int s = 16 >> i;
        const double a = pi * (double)(lane & (s - 1)) / (double)s;
        h[lr][i][lane].x = (float)cos(a);
        h[lr][i][lane].y = (float)(-sin(a));
      }
      for (int k = 0; k < 8; ++k) {
        const double a = pi * (double)(2 * lane * k) / (double)(32 * R);
        h[lr][5 + k][lane].x = k < R ? (float)cos(a) : 1.f;
        h[lr][5 + k][lane].y = k < R ? (float)(-sin(a)) : 0.f;
      }
    }
  }
  static float2 h256[256];
  for (int k = 0; k < 256; ++k) {
    h256[k].x = (float)cos(pi * (double)k / 128.0);
    h256[k].y = (float)(-sin(pi * (double)k / 128.0));
  }
  e = cudaMemcpyToSymbol(g_fft_tw, h, sizeof(h));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_fft_tw256, h256, sizeof(h256));
  if (e == cudaSuccess) done_mask |= 1ull << (dev & 63);
  return e;
}

template <int R>
struct WarpFFT {
  float2 tw_lane[5];  // W_{2s}^{lane & (s-1)}, s = 16,8,4,2,1
  float2 tw_reg[R];   // W_N^{lane * k1}
  int lane;

  __device__ __forceinline__ void init() {
    lane = threadIdx.x & 31;
    constexpr int LR = R == 1 ? 0 : (R == 2 ? 1 : (R == 4 ? 2 : 3));
#pragma unroll
    for (int i = 0; i < 5; ++i) tw_lane[i] = g_fft_tw[LR][i][lane];
#pragma unroll
    for (int k = 0; k < R; ++k) tw_reg[k] = g_fft_tw[LR][5 + k][lane];
  }

  __device__ __forceinline__ void forward(float2 (&v)[R]) {
    dft_regs<R, false>(v);
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmul(v[k], tw_reg[k]);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int s = 16 >> i;
      const bool upper = (lane & s) == 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, v[k].x, s);
        o.y = __shfl_xor_sync(0xffffffffu, v[k].y, s);
        v[k] = upper ? cadd(v[k], o) : cmul(csub(o, v[k]), tw_lane[i]);
      }
    }
  }

  // unscaled inverse: returns N * ifft
  __device__ __forceinline__ void inverse(float2 (&v)[R]) {
#pragma unroll
    for (int i = 4; i >= 0; --i) {
      const int s = 16 >> i;
      const bool upper = (lane & s) == 0;
#pragma unroll
      for (int k = 0; k < R; ++k) {
        float2 m = upper ? v[k] : cmulc(v[k], tw_lane[i]);
        float2 o;
        o.x = __shfl_xor_sync(0xffffffffu, m.x, s);
        o.y = __shfl_xor_sync(0xffffffffu, m.y, s);
        v[k] = upper ? cadd(m, o) : csub(o, m);
      }
    }
#pragma unroll
    for (int k = 1; k < R; ++k) v[k] = cmulc(v[k], tw_reg[k]);
    dft_regs<R, true>(v);
  }
};

}  // namespace tfpnp
