// Per-element bodies of the reverse-mode kernels of unet_simt.cu, written as host+device functions so the
// index arithmetic can be exercised on the CPU: tests/test_grad.py compiles this header with g++ (tests/grad_elem_host.cpp)
// and checks every function against PyTorch autograd.  The __global__ wrappers in unet_simt.cu only compute `i`.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cmath>
#include <cstring>

#ifndef __CUDACC__
inline float exp10f_portable(float x) { return powf(10.f, x); }
#define exp10f exp10f_portable
#endif
#ifdef __CUDACC__
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#define TFPNP_HD __host__ __device__ __forceinline__
namespace tfpnp { using cplx = float2; }
#else
#define TFPNP_HD inline
namespace tfpnp { struct cplx { float x, y; }; }
#endif

namespace tfpnp {

constexpr int kNumUnetConv3 = 27;
constexpr size_t kUnetParamCount = 11773857;

// UNet(2,1) 3x3 conv layer table in state_dict order (unet.py:37-46):
// {cin, cout, level} with level = log2 downsampling of the layer's resolution.
struct ConvSpec { int cin, cout, level; };
inline const ConvSpec* unet_conv_specs() {
  static const ConvSpec s[kNumUnetConv3] = {
      {2, 32, 0},    {32, 32, 0},   {32, 32, 0},     // inc
      {32, 64, 1},   {64, 64, 1},   {64, 64, 1},     // down1
      {64, 128, 2},  {128, 128, 2}, {128, 128, 2},   // down2
      {128, 256, 3}, {256, 256, 3}, {256, 256, 3},   // down3
      {256, 512, 4}, {512, 512, 4}, {512, 512, 4},   // down4
      {768, 256, 3}, {256, 256, 3}, {256, 256, 3},   // up1  (cat[skip 256, up 512])
      {384, 128, 2}, {128, 128, 2}, {128, 128, 2},   // up2  (cat[skip 128, up 256])
      {192, 64, 1},  {64, 64, 1},   {64, 64, 1},     // up3  (cat[skip 64,  up 128])
      {96, 32, 0},   {32, 32, 0},   {32, 32, 0},     // up4  (cat[skip 32,  up 64])
  };
  return s;
}

namespace grad_elem {

// d/dv LeakyReLU(0.2)(v) from the post-activation value (same sign as v; ATen uses the slope at v <= 0)
TFPNP_HD float lrelu_d(float a) { return a > 0.f ? 1.f : 0.2f; }

// g_r = gout * 1[0 <= r <= 1] (torch.clamp's backward);  gpre[b,c,p] = w[c] * g_r * lrelu'(a[b,c,p])
TFPNP_HD void outc_bwd_elem(size_t i, const float* gout, const float* r, const float* w, const float* a, float* gr,
                            float* gpre, int C, int HW) {
  const size_t b = i / HW, p = i % HW;
  const float rv = r[i];
  const float g = (rv >= 0.f && rv <= 1.f) ? gout[i] : 0.f;
  gr[i] = g;
  for (int c = 0; c < C; ++c) {
    const size_t j = (b * C + c) * HW + p;
    gpre[j] = w[c] * g * lrelu_d(a[j]);
  }
}

// Adjoint of MaxPool2d(2) (first maximum in scan order wins, as in ATen) + the skip-connection gradient + lrelu':
//   gpre[b,c,Y,X] = ((argmax(b,c,Y/2,X/2) == (Y,X) ? gpool[b,c,Y/2,X/2] : 0) + gskip[b,c,Y,X]) * lrelu'(a[b,c,Y,X])
// a, gpre: [B,C,H,W]; gpool: [B,C,H/2,W/2], i indexes it; gskip: channels [0,C) of a [B,Ccat,H,W] tensor (nullable)
TFPNP_HD void pool_bwd_elem(size_t i, const float* gpool, const float* a, const float* gskip, int Ccat, float* gpre, int C,
                            int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const int x = (int)(i % Wo), y = (int)((i / Wo) % Ho);
  const size_t bc = i / ((size_t)Wo * Ho);
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const size_t base = (bc * H + 2 * y) * W + 2 * x;
  const float v[4] = {a[base], a[base + 1], a[base + W], a[base + W + 1]};
  int arg = 0;
  float best = v[0];
  for (int k = 1; k < 4; ++k)
    if (v[k] > best) { best = v[k]; arg = k; }
  const float gp = gpool[i];
  const size_t sbase = ((b * Ccat + c) * H + 2 * y) * (size_t)W + 2 * x;
  for (int k = 0; k < 4; ++k) {
    const size_t off = (size_t)(k >> 1) * W + (k & 1);
    const float g = (k == arg ? gp : 0.f) + (gskip ? gskip[sbase + off] : 0.f);
    gpre[base + off] = g * lrelu_d(v[k]);
  }
}

// weight of high-resolution index Y on low-resolution index i for bilinear x2 with align_corners=True (the float
// expressions of upsample2_simt, so the adjoint is that of the forward kernel)
TFPNP_HD float up_weight(int Y, int i, int h, float s) {
  const float f = s * Y;
  const int y0 = (int)f;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float l = f - y0;
  return (y0 == i ? 1.f - l : 0.f) + (y1 == i ? l : 0.f);
}

// Adjoint of the x2 bilinear up-sampling + lrelu' of the low-resolution source:
//   gpre[b,c,i,j] = lrelu'(a[b,c,i,j]) * sum_{Y,X} wy(Y,i) wx(X,j) gup[b,coff+c,Y,X]
// gup: channels [coff, coff+C) of a [B,Ccat,2h,2w] tensor; a, gpre: [B,C,h,w], i indexes them.
// Rows Y in [2i-2, 2i+3] cover every Y with floor(sY) in {i-1, i} since 1/s = 2 + 1/(h-1).
TFPNP_HD void up_bwd_elem(size_t i, const float* gcat, int Ccat, int coff, const float* a, float* gpre, int C, int h,
                          int w) {
  const int xj = (int)(i % w), yi = (int)((i / w) % h);
  const size_t bc = i / ((size_t)w * h);
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const int Ho = 2 * h, Wo = 2 * w;
  const float sy = (float)(h - 1) / (float)(Ho - 1), sx = (float)(w - 1) / (float)(Wo - 1);
  const float* g = gcat + (b * Ccat + coff + c) * (size_t)Ho * Wo;
  float wx[6];
  for (int k = 0; k < 6; ++k) {
    const int X = 2 * xj - 2 + k;
    wx[k] = (X >= 0 && X < Wo) ? up_weight(X, xj, w, sx) : 0.f;
  }
  float acc = 0.f;
  for (int m = 0; m < 6; ++m) {
    const int Y = 2 * yi - 2 + m;
    if (Y < 0 || Y >= Ho) continue;
    const float wy = up_weight(Y, yi, h, sy);
    if (wy == 0.f) continue;
    float row = 0.f;
    for (int k = 0; k < 6; ++k) {
      const int X = 2 * xj - 2 + k;
      if (wx[k] != 0.f) row = fmaf(wx[k], g[(size_t)Y * Wo + X], row);
    }
    acc = fmaf(wy, row, acc);
  }
  gpre[i] = acc * lrelu_d(a[i]);
}

// weights of the input-gradient convolution: [Cout][Cin][3][3] -> [Cin][Cout][3][3] with the taps reversed
inline void transpose_flip_weights(const float* w, float* t, int cout, int cin) {
  for (int o = 0; o < cout; ++o)
    for (int c = 0; c < cin; ++c)
      for (int k = 0; k < 9; ++k) t[((size_t)c * cout + o) * 9 + (8 - k)] = w[((size_t)o * cin + c) * 9 + k];
}

// ---- the layer sequence of the denoiser's VJP, shared by the CUDA engine (UNetSimt::vjp) and the CPU emulation -------
// Workspace (floats, per image in units of HW): every layer's activation (366) | in2 2 | pooled 8 | upsampled 64 | r 1 |
// g_r 1 | gA 32 | gB 32 | gcat[level 0..3] 96 + 48 + 24 + 12.
inline size_t unet_vjp_workspace_floats(int B, int H, int W) {
  const ConvSpec* sp = unet_conv_specs();
  const size_t HW = (size_t)H * W;
  size_t units = 0;
  for (int l = 0; l < kNumUnetConv3; ++l) units += ((size_t)sp[l].cout * HW) >> (2 * sp[l].level);
  units += (2 + 8 + 64 + 1 + 1 + 32 + 32 + 96 + 48 + 24 + 12) * HW;
  return units * B;
}

// Region table of that workspace for debugging (same carve order as unet_vjp_sequence): offsets (floats) of the 27
// activations, then in2, pooled, upsampled, r, g_r, gA, gB, gcat[0..3]; 38 entries + the total as entry 38.
inline void unet_vjp_workspace_layout(int B, int H, int W, size_t out[39]) {
  const ConvSpec* sp = unet_conv_specs();
  const size_t HW = (size_t)H * W;
  size_t off = 0;
  int k = 0;
  for (int l = 0; l < kNumUnetConv3; ++l) { out[k++] = off; off += (((size_t)sp[l].cout * HW) >> (2 * sp[l].level)) * B; }
  const size_t units[11] = {2, 8, 64, 1, 1, 32, 32, 96, 48, 24, 12};
  for (int j = 0; j < 11; ++j) { out[k++] = off; off += units[j] * HW * B; }
  out[k] = off;
}

// Ops: make_input, conv (forward layer: bias + LeakyReLU), maxpool, upsample, outc_pre, outc_bwd, dgrad (input gradient
// of layer l), lrelu_bwd, pool_bwd, up_bwd, first_finish -- each returns 0 on success.  Tensors are NCHW fp32.
template <class Ops>
int unet_vjp_sequence(Ops& ops, const float* x, const float* sigma, int64_t sstride, const float* gout, float* gx,
                      float* gsigma, int64_t gs_stride, float* base, int B, int H, int W) {
  const ConvSpec* sp = unet_conv_specs();
  const size_t HW = (size_t)H * W;
  float* a[kNumUnetConv3];
  size_t off = 0;
  auto carve = [&](size_t floats_per_image) { float* p = base + off; off += floats_per_image * B; return p; };
  for (int l = 0; l < kNumUnetConv3; ++l) a[l] = carve(((size_t)sp[l].cout * HW) >> (2 * sp[l].level));
  float* in2 = carve(2 * HW);
  float* pooled = carve(8 * HW);
  float* upbuf = carve(64 * HW);
  float* r = carve(HW);
  float* gr = carve(HW);
  float* gA = carve(32 * HW);
  float* gB = carve(32 * HW);
  float* gcat[4] = {carve(96 * HW), carve(48 * HW), carve(24 * HW), carve(12 * HW)};
  const int ch[5] = {32, 64, 128, 256, 512};
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  // forward, every activation kept
  TFPNP_SEQ(ops.make_input(x, sigma, sstride, in2));
  TFPNP_SEQ(ops.conv(0, in2, 2, nullptr, 0, a[0], H, W));
  TFPNP_SEQ(ops.conv(1, a[0], 32, nullptr, 0, a[1], H, W));
  TFPNP_SEQ(ops.conv(2, a[1], 32, nullptr, 0, a[2], H, W));
  for (int lv = 1; lv <= 4; ++lv) {
    const int h = H >> lv, w = W >> lv, l0 = 3 * lv;
    TFPNP_SEQ(ops.maxpool(a[l0 - 1], pooled, ch[lv - 1], 2 * h, 2 * w));
    TFPNP_SEQ(ops.conv(l0, pooled, ch[lv - 1], nullptr, 0, a[l0], h, w));
    TFPNP_SEQ(ops.conv(l0 + 1, a[l0], ch[lv], nullptr, 0, a[l0 + 1], h, w));
    TFPNP_SEQ(ops.conv(l0 + 2, a[l0 + 1], ch[lv], nullptr, 0, a[l0 + 2], h, w));
  }
  for (int k = 0; k < 4; ++k) {
    const int lv = 3 - k, h = H >> lv, w = W >> lv, l0 = 15 + 3 * k;
    TFPNP_SEQ(ops.upsample(a[l0 - 1], upbuf, ch[lv + 1], h / 2, w / 2));
    TFPNP_SEQ(ops.conv(l0, a[3 * lv + 2], ch[lv], upbuf, ch[lv + 1], a[l0], h, w));      // cat[skip, up]  (unet.py:119)
    TFPNP_SEQ(ops.conv(l0 + 1, a[l0], ch[lv], nullptr, 0, a[l0 + 1], h, w));
    TFPNP_SEQ(ops.conv(l0 + 2, a[l0 + 1], ch[lv], nullptr, 0, a[l0 + 2], h, w));
  }
  TFPNP_SEQ(ops.outc_pre(a[26], x, r));
  // backward: `cur` = gradient w.r.t. the pre-activation output of layer l
  float* cur = gA;
  float* oth = gB;
  TFPNP_SEQ(ops.outc_bwd(gout, r, a[26], gr, cur));
  for (int l = 26; l >= 1; --l) {
    const int lv = sp[l].level, h = H >> lv, w = W >> lv;
    if (l >= 15 && (l - 15) % 3 == 0) {
      // decoder block head: input = cat[skip(level lv), up(a[l-1])]  (unet.py:99-121)
      TFPNP_SEQ(ops.dgrad(l, cur, gcat[lv], h, w));
      TFPNP_SEQ(ops.up_bwd(gcat[lv], sp[l].cin, ch[lv], a[l - 1], cur, ch[lv + 1], h / 2, w / 2));
    } else if (l <= 12 && l % 3 == 0) {
      // encoder block head: input = maxpool(a[l-1]); a[l-1] is also the skip of level lv-1  (unet.py:80-90)
      TFPNP_SEQ(ops.dgrad(l, cur, oth, h, w));
      TFPNP_SEQ(ops.pool_bwd(oth, a[l - 1], gcat[lv - 1], ch[lv - 1] + ch[lv], cur, ch[lv - 1], 2 * h, 2 * w));
    } else {
      TFPNP_SEQ(ops.dgrad(l, cur, oth, h, w));
      TFPNP_SEQ(ops.lrelu_bwd(oth, a[l - 1], (size_t)B * sp[l].cin * h * w));
      float* t = cur; cur = oth; oth = t;
    }
  }
  TFPNP_SEQ(ops.dgrad(0, cur, oth, H, W));      // [B,2,H,W]: d/dx through the network, d/d(noise map)
  TFPNP_SEQ(ops.first_finish(oth, gr, gx, gsigma, gs_stride));
#undef TFPNP_SEQ
  return 0;
}

// ---- tensor-core variant of the convolutions in the sequences above (TFPNP_GRAD_TC=1) ----------------------------------
// Activations / gradients stay fp32 NCHW (the element-wise adjoints above are unchanged); every 3x3 convolution -- forward
// recompute and input gradient -- runs on the tcgen05 kernel (unet_tc.cu: conv_v1_*) on NHWC fp16 copies:
//   fp32 NCHW --(x scale[b], round to fp16, NHWC)--> X --conv--> Y --(fp32, / scale[b], NCHW channel range)--> result
// scale[b] is a per-image power of two that puts max|g| into [512, 1024): the backward pass is linear in the cotangent, so
// the scale is exact and only guards the fp16 range.  Forward activations use scale 1.

// fp32 <-> IEEE binary16 bits, round to nearest even (device: the hardware conversion; host: the same rounding in software)
TFPNP_HD uint16_t f2h_bits(float f) {
#ifdef __CUDA_ARCH__
  return __half_as_ushort(__float2half_rn(f));
#else
  uint32_t x; memcpy(&x, &f, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | (x > 0x7f800000u ? 0x200u : 0u));   // inf / nan
  if (x >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                      // rounds to inf (>= 65520)
  if (x < 0x33000001u) return (uint16_t)sign;                                                   // rounds to zero (<= 2^-25)
  int e = (int)(x >> 23) - 127;
  uint32_t m = (x & 0x7fffffu) | 0x800000u;
  int shift = e < -14 ? 13 + (-14 - e) : 13;            // subnormal halves lose more bits
  uint32_t half_m = m >> shift;
  const uint32_t rem = m & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  if (rem > halfway || (rem == halfway && (half_m & 1u))) ++half_m;
  if (e < -14) return (uint16_t)(sign | half_m);        // subnormal (a carry into the exponent field is the right answer)
  return (uint16_t)(sign | (((uint32_t)(e + 15) << 10) + (half_m - 0x400u)));
#endif
}
TFPNP_HD float h2f_bits(uint16_t h) {
#ifdef __CUDA_ARCH__
  return __half2float(__ushort_as_half(h));
#else
  const uint32_t sign = ((uint32_t)h & 0x8000u) << 16;
  uint32_t e = (h >> 10) & 0x1fu, m = h & 0x3ffu, x;
  if (e == 0) {
    if (m == 0) x = sign;
    else { int k = 0; while (!(m & 0x400u)) { m <<= 1; ++k; } x = sign | ((uint32_t)(113 - k) << 23) | ((m & 0x3ffu) << 13); }
  } else if (e == 31) x = sign | 0x7f800000u | (m << 13);
  else x = sign | ((e + 112) << 23) | (m << 13);
  float f; memcpy(&f, &x, 4);
  return f;
#endif
}

// power-of-two scale that puts maxabs into [512, 1024); 1 for zero / non-finite input
TFPNP_HD float pow2_scale(float maxabs) {
  if (!(maxabs > 0.f) || !(maxabs < 3.0e38f)) return 1.f;
  int e;
  frexpf(maxabs, &e);                 // maxabs = m 2^e, m in [0.5, 1)
  const int k = 10 - e;
  return ldexpf(1.f, k > 100 ? 100 : k);
}

// src [B,C,HW] fp32 -> channels [coff, coff+C) of dst [B,HW,Ctot] fp16 bits, times scale[b] (nullable); i over B*C*HW.
// dst_lo (nullable): the fp16 residual plane of the split-fp16 (FP16X3) mode, v = hi + lo to ~22 bits.
// Split-fp16 operands: v = hi + lo / kLoScale with hi = fp16(v), lo = fp16((v - hi) * kLoScale).  The residual is stored
// scaled by 2^11 so that it lives in fp16's NORMAL range next to hi (an unscaled residual of a weight of magnitude 0.03 is
// 7e-6, a subnormal with 6e-8 steps: only ~7 of its 11 bits survive, which cost a factor 10 in accuracy -- measured in round 2).
// The kernels accumulate the two correction products in their own TMEM columns and scale them back in the epilogue.
constexpr float kLoScale = 2048.f;
constexpr float kLoInv = 1.f / 2048.f;

TFPNP_HD void to_half_nhwc_elem(size_t i, const float* src, uint16_t* dst, uint16_t* dst_lo, int C, int Ctot, int coff, int HW,
                                const float* scale) {
  const size_t p = i % HW, bc = i / HW;
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const float v = src[i] * (scale ? scale[b] : 1.f);
  const size_t o = (b * HW + p) * Ctot + coff + c;
  const uint16_t hi = f2h_bits(v);
  dst[o] = hi;
  if (dst_lo) dst_lo[o] = f2h_bits((v - h2f_bits(hi)) * kLoScale);
}
// src [B,HW,C] fp16 bits (+ residual plane) -> channels [coff, coff+C) of dst [B,Ctot,HW] fp32, divided by scale[b] (nullable)
TFPNP_HD void from_half_nhwc_elem(size_t i, const uint16_t* src, const uint16_t* src_lo, float* dst, int C, int Ctot, int coff,
                                  int HW, const float* scale) {
  const size_t p = i % HW, bc = i / HW;
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const float inv = scale ? 1.f / scale[b] : 1.f;      // a power of two: exact
  const size_t o = (b * HW + p) * C + c;
  const float v = h2f_bits(src[o]) + (src_lo ? h2f_bits(src_lo[o]) * kLoInv : 0.f);
  dst[(b * Ctot + coff + c) * HW + p] = v * inv;
}

// fp16 weights [tap][rows][K] of the tcgen05 convolution kernels from the state_dict tensor w [cout][cin][3][3]:
//   forward  (transpose_flip = false): rows = output channels [row0, row0+rows), K = cin
//   backward (transpose_flip = true):  rows = INPUT channels  [row0, row0+rows), K = cout, taps reversed
// out_lo (nullable): the fp16 residual plane of the split-fp16 mode
inline void build_tc_weights(const float* w, int cout, int cin, bool transpose_flip, int row0, int rows, uint16_t* out,
                             uint16_t* out_lo = nullptr) {
  for (int t = 0; t < 9; ++t)
    for (int r = 0; r < rows; ++r) {
      const int K = transpose_flip ? cout : cin;
      for (int k = 0; k < K; ++k) {
        const float v = transpose_flip ? w[((size_t)k * cin + (row0 + r)) * 9 + (8 - t)] : w[((size_t)(row0 + r) * cin + k) * 9 + t];
        const size_t o = ((size_t)t * rows + r) * K + k;
        out[o] = f2h_bits(v);
        if (out_lo) out_lo[o] = f2h_bits((v - h2f_bits(out[o])) * kLoScale);
      }
    }
}

// the output-channel parts an input-gradient convolution of layer l is split into (the tcgen05 kernel takes 32, 64 or
// multiples of 128 output channels): decoder heads split at the concatenation boundary (skip | up-sampled)
inline int dgrad_parts(int l, int rows[2]) {
  const ConvSpec& sp = unet_conv_specs()[l];
  if (l >= 15 && (l - 15) % 3 == 0) {
    const int ch[5] = {32, 64, 128, 256, 512};
    rows[0] = ch[sp.level]; rows[1] = sp.cin - rows[0];
    return 2;
  }
  rows[0] = sp.cin; rows[1] = 0;
  return 1;
}

// ---- reverse mode of torch_psnr (tfpnp/env/base.py:237-242): psnr = 10 log10(1 / mean((clamp(out,0,1) - gt)^2)) ---------
//   d psnr / d out[p] = -(10 / ln 10) (2 / HW) (clamp(out[p]) - gt[p]) 1[0 <= out[p] <= 1] / mse,   mse = 10^(-psnr / 10)
TFPNP_HD void psnr_bwd_elem(size_t i, const float* out, const float* gt, const float* psnr, const float* gpsnr, float* gout,
                            size_t HW) {
  const size_t b = i / HW;
  const float o = out[i];
  const bool inside = o >= 0.f && o <= 1.f;
  const float mse = exp10f(-0.1f * psnr[b]);
  const float k = -4.342944819032518f * 2.f / (float)HW / mse;      // -(10 / ln 10) (2 / HW) / mse
  gout[i] = inside ? gpsnr[b] * k * (o - gt[i]) : 0.f;
}

// ---- reverse mode of ADMMSolver_CSMRI.forward (csmri_variants.cu: admm_backward) -----------------------------------
// Adjoint of one iteration (x', z', u') = step(z, u; sigma, mu) with incoming (gx', gz', gu'):
//   gzt = gz' - gu';  q = ifft2c(B_mu fft2c(gzt))  (the k-space blend with y0 = 0 is self-adjoint);
//   r = ifft2c(M (fft2c(x' + u) - y0));  g_mu = <gzt, r> / (1 + mu)^2;  gxt = Re(gx' + gu' + q);
//   (gv, g_sigma) = J_D(Re(z - u), sigma)^T gxt;  gz = (gv, 0);  gu = gu' + q - (gv, 0);  gx = 0.
// States are [B,3,HW] complex (x, z, u).

// A = gz' - gu';  IN = (Re x', 0) + u   with x' = slot 0 of the next state, u = slot 2 of this state
TFPNP_HD void admm_pre_elem(size_t i, const cplx* GZ, const cplx* GU, const cplx* st_i, const cplx* st_n, cplx* A, cplx* IN,
                            int HW) {
  const size_t b = i / HW, p = i % HW;
  const cplx gz = GZ[i], gu = GU[i];
  A[i].x = gz.x - gu.x; A[i].y = gz.y - gu.y;
  const cplx u = st_i[(b * 3 + 2) * HW + p];
  const float xr = st_n[(b * 3 + 0) * HW + p].x;
  IN[i].x = xr + u.x; IN[i].y = u.y;
}
// gxt = Re(gx' + gu' + q);  GU += q;  v = Re(z - u) of this state (the denoiser input of the iteration)
TFPNP_HD void admm_mid_elem(size_t i, const cplx* GX, cplx* GU, const cplx* Q, const cplx* st_i, float* gxt, float* v,
                            int HW) {
  const size_t b = i / HW, p = i % HW;
  const cplx q = Q[i];
  cplx gu = GU[i];
  gxt[i] = GX[i].x + gu.x + q.x;
  gu.x += q.x; gu.y += q.y;
  GU[i] = gu;
  v[i] = st_i[(b * 3 + 1) * HW + p].x - st_i[(b * 3 + 2) * HW + p].x;
}
// gz = (gv, 0);  gu -= (gv, 0);  gx = 0
TFPNP_HD void admm_post_elem(size_t i, const float* gv, cplx* GX, cplx* GZ, cplx* GU) {
  const float g = gv[i];
  GX[i].x = 0.f; GX[i].y = 0.f;
  GZ[i].x = g; GZ[i].y = 0.f;
  GU[i].x -= g;
}

struct AdmmGradBufs {      // [B,HW] each
  cplx *gx, *gz, *gu, *A, *IN, *Q, *R;
  float *gxt, *v, *gv;
};

// Ops: slot_get / slot_put (slot k of a [B,3,HW] state <-> [B,HW]), pre, blend (Q = ifft2c(B_mu fft2c(A))), residual
// (R = ifft2c(M (fft2c(IN) - y0))), mu_reduce, mid, den_vjp, post -- each returns 0 on success.
// P: hyper-parameters transposed, sigma_d at P[i*B + b], mu at P[(iters + i)*B + b].  g_sigma / g_mu: [B,iters] contiguous.
template <class Ops>
int admm_backward_sequence(Ops& ops, const cplx* states, const float* P, int B, int HW, int iters, const cplx* grad_out,
                           float* g_sigma, float* g_mu, cplx* g_state_in, const AdmmGradBufs& w) {
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  TFPNP_SEQ(ops.slot_get(grad_out, w.gx, 0));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gz, 1));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gu, 2));
  const size_t state_elems = (size_t)B * HW * 3;
  const size_t np = (size_t)B * iters;
  for (int i = iters - 1; i >= 0; --i) {
    const cplx* st_i = states + (size_t)i * state_elems;
    const cplx* st_n = st_i + state_elems;
    const float* sg_i = P + (size_t)i * B;
    const float* mu_i = P + np + (size_t)i * B;
    TFPNP_SEQ(ops.pre(w.gz, w.gu, st_i, st_n, w.A, w.IN));
    TFPNP_SEQ(ops.blend(w.A, mu_i, w.Q));
    TFPNP_SEQ(ops.residual(w.IN, w.R));
    TFPNP_SEQ(ops.mu_reduce(w.A, w.R, mu_i, g_mu + i, iters));
    TFPNP_SEQ(ops.mid(w.gx, w.gu, w.Q, st_i, w.gxt, w.v));
    TFPNP_SEQ(ops.den_vjp(w.v, sg_i, w.gxt, w.gv, g_sigma + i, iters));
    TFPNP_SEQ(ops.post(w.gv, w.gx, w.gz, w.gu));
  }
  if (g_state_in) {
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gx, 0));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gz, 1));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gu, 2));
  }
#undef TFPNP_SEQ
  return 0;
}

// ---- reverse mode of ADMMSolver_SPI.forward (tasks/spi/solver.py:17-51; spi.cu: spi_backward) ---------------------------
// One iteration: z' = clamp(prox(x + u)), u' = u + x - z', x' = D(z' - u', sigma).  Under autograd the reference's
// "differentiable binary search" (transforms.py:419-437) is a constant: its iterates are midpoints of a fixed interval, so
// only the closed-form branch K1 == 0, z = (x + u) - K0 / mu, carries a gradient (through the clamp).  With incoming
// (gx', gz', gu'):   (gv, g_sigma) = J_D(z' - u', sigma)^T gx';  gzt = gz' + gv;  gut = gu' - gv;  gzt -= gut;
//   g_t = gzt 1[K1 == 0] 1[0 <= (x + u) - K0/mu <= 1];  gx = gut + g_t;  gu = gut + g_t;  gz = 0;  g_mu = sum g_t K0 / mu^2.
// States are [B,3,HW] real (x, z, u).

// v = z' - u' (the denoiser input of the iteration) from the NEXT state
TFPNP_HD void spi_v_elem(size_t i, const float* st_n, float* v, int HW) {
  const size_t b = i / HW, p = i % HW;
  v[i] = st_n[(b * 3 + 1) * HW + p] - st_n[(b * 3 + 2) * HW + p];
}
// in place on (GX, GZ, GU); term[i] = this pixel's contribution to g_mu
TFPNP_HD void spi_step_elem(size_t i, const float* st_i, const float* x0, const float* K, int64_t K_stride, const float* mu,
                            const float* gv, float* GX, float* GZ, float* GU, float* term, int HW) {
  const size_t b = i / HW, p = i % HW;
  const float K10 = K[b * K_stride] * 10.f;                // solver.py:32
  const float Ksq = K10 * K10;
  const float K1 = x0[i] * Ksq;                            // solver.py:33
  const float m = mu[b];
  const float g = gv[i];
  const float gut = GU[i] - g;
  const float gzt = GZ[i] + g - gut;
  float gt = 0.f, t = 0.f;
  if (K1 == 0.f) {
    const float K0 = Ksq - K1;
    const float zpre = (st_i[(b * 3 + 0) * HW + p] + st_i[(b * 3 + 2) * HW + p]) - K0 / m;
    if (zpre >= 0.f && zpre <= 1.f) { gt = gzt; t = gzt * K0 / (m * m); }
  }
  GX[i] = gut + gt;
  GU[i] = gut + gt;
  GZ[i] = 0.f;
  term[i] = t;
}

struct SpiGradBufs { float *gx, *gz, *gu, *v, *gv, *term; };   // [B,HW] each

// Ops: slot_get / slot_put, make_v, den_vjp(v, sigma_i, gx', gv, g_sigma ptr, stride), step, reduce(term, g_mu ptr, stride)
template <class Ops>
int spi_backward_sequence(Ops& ops, const float* states, const float* P, int B, int HW, int iters, const float* grad_out,
                          float* g_sigma, float* g_mu, float* g_state_in, const SpiGradBufs& w) {
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  TFPNP_SEQ(ops.slot_get(grad_out, w.gx, 0));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gz, 1));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gu, 2));
  const size_t state_elems = (size_t)B * HW * 3;
  const size_t np = (size_t)B * iters;
  for (int i = iters - 1; i >= 0; --i) {
    const float* st_i = states + (size_t)i * state_elems;
    const float* st_n = st_i + state_elems;
    const float* sg_i = P + (size_t)i * B;
    const float* mu_i = P + np + (size_t)i * B;
    TFPNP_SEQ(ops.make_v(st_n, w.v));
    TFPNP_SEQ(ops.den_vjp(w.v, sg_i, w.gx, w.gv, g_sigma + i, iters));
    TFPNP_SEQ(ops.step(st_i, mu_i, w.gv, w.gx, w.gz, w.gu, w.term));
    TFPNP_SEQ(ops.reduce(w.term, g_mu + i, iters));
  }
  if (g_state_in) {
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gx, 0));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gz, 1));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gu, 2));
  }
#undef TFPNP_SEQ
  return 0;
}

// ---- reverse mode of IADMMSolver_CT.forward (tasks/ct/solver.py:17-53; misc.cu: ct_backward) ----------------------------
// One iteration: x' = D(z - u, sigma);  z' = z - tau (bp(z) + mu r),  bp(z) = A^T(A z - y0)/opnorm^2,  r = z - x' - u;
// u' = u + x' - z'.  With incoming (gx', gz', gu'):  gxt = gx' + gu';  gzt = gz' - gu';
//   gz = gzt - tau (A^T A gzt / opnorm^2 + mu gzt);  gxt += tau mu gzt;  gu = gu' + tau mu gzt;
//   g_tau = -<gzt, bp(z) + mu r>;  g_mu = -tau <gzt, r>;  (gv, g_sigma) = J_D(z - u, sigma)^T gxt;  gz += gv;  gu -= gv;  gx = 0.
// A^T A is symmetric because the backprojector is the exact transpose of the projector.  States are [B,3,HW] real.

// gzt = gz' - gu' -> GZT;  z of this state -> Z (contiguous operands for the projector)
TFPNP_HD void ct_pre_elem(size_t i, const float* GZ, const float* GU, const float* st_i, float* GZT, float* Z, int HW) {
  const size_t b = i / HW, p = i % HW;
  GZT[i] = GZ[i] - GU[i];
  Z[i] = st_i[(b * 3 + 1) * HW + p];
}
// W1 = A^T A gzt, W2 = A^T (A z - y0) (both un-normalised).  In place on (GX -> gxt, GZ, GU); v = z - u;
// term_tau / term_mu: this pixel's contributions to g_tau / g_mu
TFPNP_HD void ct_mid_elem(size_t i, const float* st_i, const float* st_n, const float* GZT, const float* W1, const float* W2,
                          const float* mu, const float* tau, float inv_opnorm2, float* GX, float* GZ, float* GU, float* v,
                          float* term_tau, float* term_mu, int HW) {
  const size_t b = i / HW, p = i % HW;
  const float m = mu[b], t = tau[b];
  const float z = st_i[(b * 3 + 1) * HW + p], u = st_i[(b * 3 + 2) * HW + p], xn = st_n[(b * 3 + 0) * HW + p];
  const float r = z - xn - u;
  const float gzt = GZT[i];
  const float gu_in = GU[i];
  GZ[i] = gzt - t * (W1[i] * inv_opnorm2 + m * gzt);
  GX[i] = GX[i] + gu_in + t * m * gzt;          // gxt
  GU[i] = gu_in + t * m * gzt;
  term_tau[i] = -gzt * (W2[i] * inv_opnorm2 + m * r);
  term_mu[i] = -t * gzt * r;
  v[i] = z - u;
}
// gz += gv;  gu -= gv;  gx = 0
TFPNP_HD void ct_post_elem(size_t i, const float* gv, float* GX, float* GZ, float* GU) {
  const float g = gv[i];
  GX[i] = 0.f;
  GZ[i] += g;
  GU[i] -= g;
}

struct CtGradBufs { float *gx, *gz, *gu, *gzt, *z, *w1, *w2, *v, *gv, *t_tau, *t_mu; };   // [B,HW] each

// Ops: slot_get / slot_put, pre, ata (W = A^T (A img - y0?) un-normalised; with_y0 selects the residual form), mid, reduce
// (per-image sum -> strided output), den_vjp, post.  P: [sigma | mu | tau][iters][B].
template <class Ops>
int ct_backward_sequence(Ops& ops, const float* states, const float* P, int B, int HW, int iters, const float* grad_out,
                         float* g_sigma, float* g_mu, float* g_tau, float* g_state_in, const CtGradBufs& w) {
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  TFPNP_SEQ(ops.slot_get(grad_out, w.gx, 0));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gz, 1));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gu, 2));
  const size_t state_elems = (size_t)B * HW * 3;
  const size_t np = (size_t)B * iters;
  for (int i = iters - 1; i >= 0; --i) {
    const float* st_i = states + (size_t)i * state_elems;
    const float* st_n = st_i + state_elems;
    const float* sg_i = P + (size_t)i * B;
    const float* mu_i = P + np + (size_t)i * B;
    const float* tau_i = P + 2 * np + (size_t)i * B;
    TFPNP_SEQ(ops.pre(w.gz, w.gu, st_i, w.gzt, w.z));
    TFPNP_SEQ(ops.ata(w.gzt, false, w.w1));
    TFPNP_SEQ(ops.ata(w.z, true, w.w2));
    TFPNP_SEQ(ops.mid(st_i, st_n, w.gzt, w.w1, w.w2, mu_i, tau_i, w.gx, w.gz, w.gu, w.v, w.t_tau, w.t_mu));
    TFPNP_SEQ(ops.reduce(w.t_tau, g_tau + i, iters));
    TFPNP_SEQ(ops.reduce(w.t_mu, g_mu + i, iters));
    TFPNP_SEQ(ops.den_vjp(w.v, sg_i, w.gx, w.gv, g_sigma + i, iters));
    TFPNP_SEQ(ops.post(w.gv, w.gx, w.gz, w.gu));
  }
  if (g_state_in) {
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gx, 0));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gz, 1));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gu, 2));
  }
#undef TFPNP_SEQ
  return 0;
}

// ---- reverse mode of IADMMSolver_PR.forward (tasks/pr/solver.py:37-76; pr.cu: pr_backward) ------------------------------
// One iteration: x' = (D(Re(z - u), sigma), 0);  z' = z - tau (g(z) + mu r),  r = z - x' - u;  u' = u + x' - z', with
//   g(z) = mean_j conj(m_j) . IFFT(h(FFT(m_j . z))),   h(w) = (1 - y0/|w|) w          (cdp_forward / cdp_backward, ortho FFTs).
// h is a map R^2 -> R^2 with the SYMMETRIC Jacobian J_h(w) b = b - y0 (b/|w| - w (w.b)/|w|^3), the FFT is unitary and the
// adjoint of multiplying by conj(m) is multiplying by m, so  J_g^T c = mean_j conj(m_j) . IFFT(J_h(w_j) FFT(m_j . c)):
// the forward operator with h replaced by its Jacobian.  With incoming (gx', gz', gu'):  gzt = gz' - gu';
//   gz = gzt - tau (J_g^T gzt + mu gzt);  gxt = gx' + gu' + tau mu gzt;  gu = gu' + tau mu gzt;
//   g_tau = -<gzt, g(z) + mu r>;  g_mu = -tau <gzt, r>;  (gv, g_sigma) = J_D^T Re(gxt);  gz += (gv,0);  gu -= (gv,0);  gx = 0.
// States are [B,3,HW] complex; mask [B,M,HW] complex; y0 [B,M,HW] real (natural FFT order).

TFPNP_HD cplx c_mul(cplx a, cplx b) { cplx r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
TFPNP_HD cplx c_mul_conj(cplx a, cplx m) { cplx r; r.x = a.x * m.x + a.y * m.y; r.y = a.y * m.x - a.x * m.y; return r; }   // a conj(m)

// out[b,j,p] = mask[b,j,p] img[b,p];  i over B*M*HW
TFPNP_HD void pr_mul_elem(size_t i, const cplx* img, const cplx* mask, cplx* out, int M, int HW) {
  const size_t p = i % HW, b = i / ((size_t)M * HW);
  out[i] = c_mul(img[b * HW + p], mask[i]);
}
// in place: W = FFT(m z) -> h(W);  Bc = FFT(m c) -> J_h(W) Bc;  i over B*M*HW
TFPNP_HD void pr_h_elem(size_t i, cplx* W, cplx* Bc, const float* y0) {
  const cplx w = W[i], b = Bc[i];
  const float a = sqrtf(w.x * w.x + w.y * w.y);
  const float y = y0[i];
  const float ratio = (a - y) / a;                       // unguarded like the reference (solver.py:67)
  W[i].x = ratio * w.x; W[i].y = ratio * w.y;
  const float dot = w.x * b.x + w.y * b.y;
  const float k = y * dot / (a * a * a);
  Bc[i].x = b.x - y * b.x / a + k * w.x;
  Bc[i].y = b.y - y * b.y / a + k * w.y;
}
// out[b,p] = (1/M) sum_j E[b,j,p] conj(mask[b,j,p]);  i over B*HW
TFPNP_HD void pr_acc_elem(size_t i, const cplx* E, const cplx* mask, cplx* out, int M, int HW) {
  const size_t p = i % HW, b = i / HW;
  float sx = 0.f, sy = 0.f;
  for (int j = 0; j < M; ++j) {
    const size_t o = (b * M + j) * HW + p;
    const cplx t = c_mul_conj(E[o], mask[o]);
    sx += t.x; sy += t.y;
  }
  out[i].x = sx / (float)M; out[i].y = sy / (float)M;
}
TFPNP_HD void pr_pre_elem(size_t i, const cplx* GZ, const cplx* GU, const cplx* st_i, cplx* GZT, cplx* Z, int HW) {
  const size_t b = i / HW, p = i % HW;
  GZT[i].x = GZ[i].x - GU[i].x; GZT[i].y = GZ[i].y - GU[i].y;
  Z[i] = st_i[(b * 3 + 1) * HW + p];
}
// Gz = g(z), JC = J_g^T gzt.  In place on (GZ, GU); gxt_re = Re(gx' + gu' + tau mu gzt); v = Re(z - u)
TFPNP_HD void pr_mid_elem(size_t i, const cplx* st_i, const cplx* st_n, const cplx* GZT, const cplx* JC, const cplx* Gz,
                          const float* mu, const float* tau, const cplx* GX, cplx* GZ, cplx* GU, float* gxt_re, float* v,
                          float* term_tau, float* term_mu, int HW) {
  const size_t b = i / HW, p = i % HW;
  const float m = mu[b], t = tau[b];
  const cplx z = st_i[(b * 3 + 1) * HW + p], u = st_i[(b * 3 + 2) * HW + p], xn = st_n[(b * 3 + 0) * HW + p];
  const float rx = z.x - xn.x - u.x, ry = z.y - xn.y - u.y;
  const cplx gzt = GZT[i], jc = JC[i], g = Gz[i];
  cplx gu = GU[i];
  GZ[i].x = gzt.x - t * (jc.x + m * gzt.x);
  GZ[i].y = gzt.y - t * (jc.y + m * gzt.y);
  gxt_re[i] = GX[i].x + gu.x + t * m * gzt.x;
  gu.x += t * m * gzt.x; gu.y += t * m * gzt.y;
  GU[i] = gu;
  term_tau[i] = -(gzt.x * (g.x + m * rx) + gzt.y * (g.y + m * ry));
  term_mu[i] = -t * (gzt.x * rx + gzt.y * ry);
  v[i] = z.x - u.x;
}

struct PrGradBufs {
  cplx *gx, *gz, *gu, *gzt, *z, *gzv, *jc;     // [B,HW]
  cplx *w, *w2, *bc, *bc2;                      // [B,M,HW]
  float *gxt, *v, *gv, *t_tau, *t_mu;           // [B,HW]
};

// Ops: slot_get / slot_put (complex), pre, mul, fft(in, out, inverse) on B*M images (un-centred, ortho), h, acc, mid, reduce,
// den_vjp, post (= admm_post_elem).  P: [sigma | mu | tau][iters][B].
template <class Ops>
int pr_backward_sequence(Ops& ops, const cplx* states, const float* P, int B, int HW, int iters, const cplx* grad_out,
                         float* g_sigma, float* g_mu, float* g_tau, cplx* g_state_in, const PrGradBufs& w) {
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  TFPNP_SEQ(ops.slot_get(grad_out, w.gx, 0));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gz, 1));
  TFPNP_SEQ(ops.slot_get(grad_out, w.gu, 2));
  const size_t state_elems = (size_t)B * HW * 3;
  const size_t np = (size_t)B * iters;
  for (int i = iters - 1; i >= 0; --i) {
    const cplx* st_i = states + (size_t)i * state_elems;
    const cplx* st_n = st_i + state_elems;
    const float* sg_i = P + (size_t)i * B;
    const float* mu_i = P + np + (size_t)i * B;
    const float* tau_i = P + 2 * np + (size_t)i * B;
    TFPNP_SEQ(ops.pre(w.gz, w.gu, st_i, w.gzt, w.z));
    TFPNP_SEQ(ops.mul(w.z, w.w2));                 // m_j z
    TFPNP_SEQ(ops.fft(w.w2, w.w, false));          // W = FFT(m_j z)
    TFPNP_SEQ(ops.mul(w.gzt, w.bc2));              // m_j gzt
    TFPNP_SEQ(ops.fft(w.bc2, w.bc, false));        // Bc = FFT(m_j gzt)
    TFPNP_SEQ(ops.h(w.w, w.bc));                   // W = h(W), Bc = J_h(W) Bc
    TFPNP_SEQ(ops.fft(w.w, w.w2, true));
    TFPNP_SEQ(ops.acc(w.w2, w.gzv));               // g(z)
    TFPNP_SEQ(ops.fft(w.bc, w.bc2, true));
    TFPNP_SEQ(ops.acc(w.bc2, w.jc));               // J_g^T gzt
    TFPNP_SEQ(ops.mid(st_i, st_n, w.gzt, w.jc, w.gzv, mu_i, tau_i, w.gx, w.gz, w.gu, w.gxt, w.v, w.t_tau, w.t_mu));
    TFPNP_SEQ(ops.reduce(w.t_tau, g_tau + i, iters));
    TFPNP_SEQ(ops.reduce(w.t_mu, g_mu + i, iters));
    TFPNP_SEQ(ops.den_vjp(w.v, sg_i, w.gxt, w.gv, g_sigma + i, iters));
    TFPNP_SEQ(ops.post(w.gv, w.gx, w.gz, w.gu));    // gz.re += gv ... see pr_post_elem
  }
  if (g_state_in) {
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gx, 0));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gz, 1));
    TFPNP_SEQ(ops.slot_put(g_state_in, w.gu, 2));
  }
#undef TFPNP_SEQ
  return 0;
}
// gz += (gv, 0);  gu -= (gv, 0);  gx = 0
TFPNP_HD void pr_post_elem(size_t i, const float* gv, cplx* GX, cplx* GZ, cplx* GU) {
  const float g = gv[i];
  GX[i].x = 0.f; GX[i].y = 0.f;
  GZ[i].x += g;
  GU[i].x -= g;
}

// ---- reverse mode of the other CS-MRI solvers (tasks/csmri/solver.py:60-201; csmri_variants.cu: variant_backward) --------
// Building blocks: the k-space blend with y0 = 0 and the masked projection F^-1 M F are self-adjoint (unitary F, real diagonal),
// R(w) = ifft2c(M (fft2c(w) - y0)) is the masked residual.  algo: 1 HQS (x,z | sigma,mu), 2 PG (x | sigma,tau),
// 3 APG (x_prev,s | sigma,tau,beta), 4 RED-ADMM (x,z,u | sigma,mu,lamda).  G0..G2: running cotangents of the state slots.
//   HQS:  x' = D(Re z); z' = Blend_mu(x').            gxt = Re(gx' + Blend0(gz')); g_mu = <gz', R(x')>/(1+mu)^2; gz = (gv,0); gx = 0
//   PG:   z = x - tau R(x); x' = D(Re z).             a = (gv,0); gx = a - tau P(a); g_tau = -<a, R(x)>          (P = F^-1 M F)
//   APG:  z = s - tau R(s); x' = D(Re z); s' = x' + beta (x' - x).   gxt = Re(gx' + (1+beta) gs'); g_beta = <gs', x' - x>;
//         gx = -beta gs'; a = (gv,0); gs = a - tau P(a); g_tau = -<a, R(s)>
//   RED:  xh = D(Re x); x' = (lam xh + mu (z-u))/(mu+lam); z' = Blend_mu(x'+u); u' = u + x' - z'.   gzt = gz' - gu';
//         q = Blend0(gzt); gx't = gx' + gu' + q; gu = gu' + q - mu/(mu+lam) gx't; gz = mu/(mu+lam) gx't; gxh = lam/(mu+lam) Re gx't;
//         g_mu = <gzt, R(x'+u)>/(1+mu)^2 + <gx't, (z-u) - x'>/(mu+lam); g_lam = <gx't, xh - x'>/(mu+lam); gx = (gv,0)
TFPNP_HD float c_dot(cplx a, cplx b) { return a.x * b.x + a.y * b.y; }

TFPNP_HD void var_pre_elem(int algo, size_t i, const cplx* st_i, const cplx* st_n, const cplx* G1, const cplx* G2, cplx* A, cplx* IN,
                           int V, int HW) {
  const size_t b = i / HW, p = i % HW;
  const cplx* si = st_i + (b * V) * HW + p;
  const cplx* sn = st_n + (b * V) * HW + p;
  if (algo == 1) { A[i] = G1[i]; IN[i] = sn[0]; }
  else if (algo == 2) { IN[i] = si[0]; }
  else if (algo == 3) { IN[i] = si[(size_t)HW]; }
  else { A[i].x = G1[i].x - G2[i].x; A[i].y = G1[i].y - G2[i].y;
         IN[i].x = sn[0].x + si[2 * (size_t)HW].x; IN[i].y = sn[0].y + si[2 * (size_t)HW].y; }
}
// after the FFT steps: Q = Blend0(A) (HQS, RED), R = R(IN).  Writes gxt, v (the denoiser VJP's cotangent / input), the
// per-pixel terms of d/dp1, d/dp2, and updates the cotangents that do not depend on gv.
TFPNP_HD void var_mid_elem(int algo, size_t i, const cplx* st_i, const cplx* st_n, const cplx* A, const cplx* IN, const cplx* Q,
                           const cplx* R, const float* p1, const float* p2, cplx* G0, cplx* G1, cplx* G2, float* gxt, float* v,
                           float* t1, float* t2, int V, int HW) {
  const size_t b = i / HW, p = i % HW;
  const cplx* si = st_i + (b * V) * HW + p;
  const cplx* sn = st_n + (b * V) * HW + p;
  t1[i] = 0.f; t2[i] = 0.f;
  if (algo == 1) {
    const float m = 1.f + p1[b];
    t1[i] = c_dot(A[i], R[i]) / (m * m);
    gxt[i] = G0[i].x + Q[i].x;
    v[i] = si[(size_t)HW].x;
  } else if (algo == 2) {
    v[i] = IN[i].x - p1[b] * R[i].x;
    gxt[i] = G0[i].x;
  } else if (algo == 3) {
    const float bt = p2[b];
    const cplx gs = G1[i];
    v[i] = IN[i].x - p1[b] * R[i].x;
    gxt[i] = G0[i].x + (1.f + bt) * gs.x;
    cplx d; d.x = sn[0].x - si[0].x; d.y = sn[0].y - si[0].y;
    t2[i] = c_dot(gs, d);
    G0[i].x = -bt * gs.x; G0[i].y = -bt * gs.y;
  } else {
    const float mu = p1[b], lam = p2[b], k = 1.f / (mu + lam), m = 1.f + mu;
    const cplx q = Q[i], xn = sn[0], z = si[(size_t)HW], u = si[2 * (size_t)HW];
    cplx g; g.x = G0[i].x + G2[i].x + q.x; g.y = G0[i].y + G2[i].y + q.y;          // cotangent of x'
    cplx zu; zu.x = z.x - u.x; zu.y = z.y - u.y;
    cplx xh; xh.x = ((mu + lam) * xn.x - mu * zu.x) / lam; xh.y = 0.f;               // x_half recovered from x'
    cplx dl; dl.x = xh.x - xn.x; dl.y = xh.y - xn.y;
    cplx dm; dm.x = zu.x - xn.x; dm.y = zu.y - xn.y;
    t1[i] = c_dot(A[i], R[i]) / (m * m) + c_dot(g, dm) * k;
    t2[i] = c_dot(g, dl) * k;
    gxt[i] = lam * k * g.x;
    G1[i].x = mu * k * g.x; G1[i].y = mu * k * g.y;
    G2[i].x = G2[i].x + q.x - mu * k * g.x; G2[i].y = G2[i].y + q.y - mu * k * g.y;
    v[i] = si[0].x;
  }
}
// after the denoiser VJP: HQS / RED finish here; PG / APG set A = (gv, 0) for the projection step
TFPNP_HD void var_post1_elem(int algo, size_t i, const float* gv, cplx* A, cplx* G0, cplx* G1) {
  const float g = gv[i];
  if (algo == 1) { G0[i].x = 0.f; G0[i].y = 0.f; G1[i].x = g; G1[i].y = 0.f; }
  else if (algo == 4) { G0[i].x = g; G0[i].y = 0.f; }
  else { A[i].x = g; A[i].y = 0.f; }
}
// PG / APG: Q = P(A);  g = A - tau Q;  t1 = -<A, R>
TFPNP_HD void var_post2_elem(int algo, size_t i, const cplx* A, const cplx* Q, const cplx* R, const float* p1, cplx* G0, cplx* G1,
                             float* t1, int HW) {
  const size_t b = i / HW;
  const float tau = p1[b];
  cplx g; g.x = A[i].x - tau * Q[i].x; g.y = A[i].y - tau * Q[i].y;
  t1[i] = -c_dot(A[i], R[i]);
  if (algo == 2) G0[i] = g; else G1[i] = g;
}

struct VarGradBufs { cplx *g0, *g1, *g2, *A, *IN, *Q, *R; float *gxt, *v, *gv, *t1, *t2; };

// Ops: slot_get / slot_put (V slots), pre, blend0(A, mu, Q), resid(in, with_y0, out), mid, den_vjp, post1, post2, reduce.
// P: [p0 | p1 | p2][iters][B].  g_p2 may be null for the two-parameter solvers.
template <class Ops>
int variant_backward_sequence(Ops& ops, int algo, const cplx* states, const float* P, int B, int HW, int iters,
                              const cplx* grad_out, float* g_p0, float* g_p1, float* g_p2, cplx* g_state_in,
                              const VarGradBufs& w) {
#define TFPNP_SEQ(expr) do { int _s = (expr); if (_s != 0) return _s; } while (0)
  const int V = algo == 2 ? 1 : (algo == 4 ? 3 : 2);
  cplx* G[3] = {w.g0, w.g1, w.g2};
  for (int k = 0; k < V; ++k) TFPNP_SEQ(ops.slot_get(grad_out, G[k], V, k));
  const size_t state_elems = (size_t)B * HW * V;
  const size_t np = (size_t)B * iters;
  for (int i = iters - 1; i >= 0; --i) {
    const cplx* st_i = states + (size_t)i * state_elems;
    const cplx* st_n = st_i + state_elems;
    const float* p0 = P + (size_t)i * B;
    const float* p1 = P + np + (size_t)i * B;
    const float* p2 = P + 2 * np + (size_t)i * B;
    TFPNP_SEQ(ops.pre(algo, st_i, st_n, w.g1, w.g2, w.A, w.IN, V));
    if (algo == 1 || algo == 4) TFPNP_SEQ(ops.blend0(w.A, p1, w.Q));
    TFPNP_SEQ(ops.resid(w.IN, true, w.R));
    TFPNP_SEQ(ops.mid(algo, st_i, st_n, w.A, w.IN, w.Q, w.R, p1, p2, w.g0, w.g1, w.g2, w.gxt, w.v, w.t1, w.t2, V));
    if (algo == 1 || algo == 4) TFPNP_SEQ(ops.reduce(w.t1, g_p1 + i, iters));
    if (algo >= 3) TFPNP_SEQ(ops.reduce(w.t2, g_p2 + i, iters));
    TFPNP_SEQ(ops.den_vjp(w.v, p0, w.gxt, w.gv, g_p0 + i, iters));
    TFPNP_SEQ(ops.post1(algo, w.gv, w.A, w.g0, w.g1));
    if (algo == 2 || algo == 3) {
      TFPNP_SEQ(ops.resid(w.A, false, w.Q));
      TFPNP_SEQ(ops.post2(algo, w.A, w.Q, w.R, p1, w.g0, w.g1, w.t1));
      TFPNP_SEQ(ops.reduce(w.t1, g_p1 + i, iters));
    }
  }
  if (g_state_in)
    for (int k = 0; k < V; ++k) TFPNP_SEQ(ops.slot_put(g_state_in, G[k], V, k));
#undef TFPNP_SEQ
  return 0;
}

}  // namespace grad_elem
}  // namespace tfpnp
