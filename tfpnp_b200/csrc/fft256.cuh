// 256-point complex FFT over a half-warp: 16 lanes x 16 points, lane t keeps v[j] = x[t + 16 j].
//
//   n = t + 16 j,  k = k1 + 16 k2:   X[k1 + 16 k2] = sum_t W_16^{t k2} ( W_256^{t k1} sum_j x[t + 16 j] W_16^{j k1} )
//
// i.e. a 16-point DFT in registers (two radix-4 passes), a lane twiddle, ONE exchange through shared memory (lane t
// hands its k1-th value to lane k1) and a second 16-point DFT in registers.  The result comes back in NATURAL order in
// the same distribution (lane k1 keeps X[k1 + 16 k2] in v[k2]), so forward and inverse are the same routine with
// conjugated twiddles and frequency-domain operands need no permutation.
//
// Why next to WarpFFT (fft.cuh): at N = 256 the warp-wide transform spends 5 butterfly stages x 8 registers x
// (2 SHFL + both branches of the butterfly) ~ 600 instructions per transform and the PR kernels were issue-bound
// (pr_cols: 67 M warp instructions, 51 % issue-active, 109 us -- profiles/r02_ncu_upd_pr.txt).  Sixteen points per lane
// need ~480 instructions per lane for TWO transforms per warp: 2.5x fewer issue slots per transform.
//
// Exchange slots: slot(a, b) = base[(17 a + b) * stride].  Lane t writes slot(k1, t) and reads slot(t, t'): with a row
// of 17 both directions touch 16 different bank pairs (stride 1: 2t and 34t mod 32; the column tiles of pr256_cols use
// the same slots with stride = tile pitch).  A tile column that stores element n at row n + (n >> 4) IS slot(j, t) for
// n = t + 16 j, so the column kernel transforms in place without a separate exchange buffer.
#pragma once
#include "fft.cuh"

#ifdef __CUDA_ARCH__
#define TFPNP_UNROLL _Pragma("unroll")
#else
#define TFPNP_UNROLL   // (the host pass of the __host__ __device__ bodies: gcc does not know the pragma)
#endif

namespace tfpnp {

template <bool INV>
__host__ __device__ __forceinline__ float2 tw_mul(float2 a, float wr, float wi) {   // a * (wr + i wi), conjugated for INV
  if (INV) wi = -wi;
  return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
}

// natural order in, natural order out; INV: conjugated twiddles, unscaled
template <bool INV>
__host__ __device__ __forceinline__ void dft16(float2 (&v)[16]) {
  const float c = 0.92387953251128675613f, s = 0.38268343236508977173f, h = 0.70710678118654752440f;
  // j = j1 + 4 j2, k = k2 + 4 k1:  W16^{jk} = W16^{j1 k2} W4^{j1 k1} W4^{j2 k2}
TFPNP_UNROLL
  for (int j1 = 0; j1 < 4; ++j1) dft4<INV>(v[j1], v[j1 + 4], v[j1 + 8], v[j1 + 12]);   // v[j1 + 4 k2] = Z[j1][k2]
  v[5] = tw_mul<INV>(v[5], c, -s);                                   // W16^1
  v[9] = tw_mul<INV>(v[9], h, -h);                                   // W16^2
  v[13] = tw_mul<INV>(v[13], s, -c);                                 // W16^3
  v[6] = tw_mul<INV>(v[6], h, -h);                                   // W16^2
  v[10] = mul_mi<INV>(v[10]);                                        // W16^4 = -i
  v[14] = tw_mul<INV>(v[14], -h, -h);                                // W16^6
  v[7] = tw_mul<INV>(v[7], s, -c);                                   // W16^3
  v[11] = tw_mul<INV>(v[11], -h, -h);                                // W16^6
  v[15] = tw_mul<INV>(v[15], -c, s);                                 // W16^9
TFPNP_UNROLL
  for (int k2 = 0; k2 < 4; ++k2) dft4<INV>(v[4 * k2], v[4 * k2 + 1], v[4 * k2 + 2], v[4 * k2 + 3]);   // v[4 k2 + k1] = X[k2 + 4 k1]
  float2 x;   // 4x4 transpose of the register file: pure renaming once unrolled
  x = v[1]; v[1] = v[4]; v[4] = x;
  x = v[2]; v[2] = v[8]; v[8] = x;
  x = v[3]; v[3] = v[12]; v[12] = x;
  x = v[6]; v[6] = v[9]; v[9] = x;
  x = v[7]; v[7] = v[13]; v[13] = x;
  x = v[11]; v[11] = v[14]; v[14] = x;
}

constexpr int kF256Slots = 16 * 17;   // float2 slots of one half-warp's exchange area (stride 1)

// The stage before the exchange and the stage after it, split so that a host test can run the sixteen lanes in turn.
template <bool INV>
__host__ __device__ __forceinline__ void fft256_pre(float2 (&v)[16], int t, const float2* tw256) {
  dft16<INV>(v);
TFPNP_UNROLL
  for (int k1 = 1; k1 < 16; ++k1) {
    const float2 w = tw256[(t * k1) & 255];
    v[k1] = INV ? cmulc(v[k1], w) : cmul(v[k1], w);
  }
}

#ifdef __CUDACC__
// One transform per half-warp; all 32 lanes of the warp must call it together (two transforms side by side).
// base/stride: this half-warp's slots (see the header comment); tw256: the 256-entry table in shared memory.
template <bool INV, int STRIDE>
__device__ __forceinline__ void fft256_run(float2 (&v)[16], int t, float2* base, const float2* tw256) {
  fft256_pre<INV>(v, t, tw256);
  __syncwarp();   // every lane has finished reading its previous values out of the slots
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) base[(17 * k1 + t) * STRIDE] = v[k1];
  __syncwarp();
#pragma unroll
  for (int tt = 0; tt < 16; ++tt) v[tt] = base[(17 * t + tt) * STRIDE];
  dft16<INV>(v);
}
#endif

}  // namespace tfpnp
