// UNet(2,1) denoiser on CUDA cores in fp32 (TFPNP_PREC_FP32_SIMT).
//
// Verification mode: the same network as tfpnp/pnp/denoiser/models/unet.py:34-131
// evaluated with plain FFMA so the tensor-core path (unet_tc.cu) can be checked
// on the GPU against an fp32 result that shares nothing with it but the weights.
// NCHW fp32 activations, direct 3x3 convolution with shared-memory halo tiles.
#include "common.cuh"
#include <vector>

namespace tfpnp {
namespace {

constexpr int TS = 16;    // output tile edge
constexpr int CI_T = 8;   // input channels staged per step
constexpr int CO_T = 16;  // output channels per block

// out[b,co,y,x] = act(bias[co] + sum_{ci,ky,kx} w[co,ci,ky,kx] * in[b,ci,y+ky-1,x+kx-1])
// the input is the channel-concatenation of src0 (C0 ch) and src1 (C1 ch)  (unet.py:119)
__global__ void __launch_bounds__(TS* TS)
conv3x3_simt(const float* __restrict__ src0, int C0, const float* __restrict__ src1, int C1,
             const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
             int Cout, int H, int W, int leaky) {
  __shared__ float s_in[CI_T][TS + 2][TS + 2];
  __shared__ float s_w[CI_T][9][CO_T];
  const int tiles_x = (W + TS - 1) / TS;
  const int tx0 = (blockIdx.x % tiles_x) * TS, ty0 = (blockIdx.x / tiles_x) * TS;
  const int co0 = blockIdx.y * CO_T;
  const int b = blockIdx.z;
  const int tx = threadIdx.x % TS, ty = threadIdx.x / TS;
  const int Cin = C0 + C1;
  float acc[CO_T];
#pragma unroll
  for (int i = 0; i < CO_T; ++i) acc[i] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += CI_T) {
    for (int i = threadIdx.x; i < CI_T * (TS + 2) * (TS + 2); i += TS * TS) {
      int ci = i / ((TS + 2) * (TS + 2));
      int r = i % ((TS + 2) * (TS + 2));
      int yy = ty0 + r / (TS + 2) - 1, xx = tx0 + r % (TS + 2) - 1;
      int c = c0 + ci;
      float v = 0.f;
      if (c < Cin && yy >= 0 && yy < H && xx >= 0 && xx < W) {
        v = (c < C0) ? src0[((size_t)(b * C0 + c) * H + yy) * W + xx]
                     : src1[((size_t)(b * C1 + (c - C0)) * H + yy) * W + xx];
      }
      s_in[ci][r / (TS + 2)][r % (TS + 2)] = v;
    }
    for (int i = threadIdx.x; i < CI_T * 9 * CO_T; i += TS * TS) {
      int co = i % CO_T, k = (i / CO_T) % 9, ci = i / (CO_T * 9);
      int c = c0 + ci;
      float v = 0.f;
      if (c < Cin && co0 + co < Cout) v = w[((size_t)(co0 + co) * Cin + c) * 9 + k];
      s_w[ci][k][co] = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int ci = 0; ci < CI_T; ++ci) {
      float v[9];
#pragma unroll
      for (int k = 0; k < 9; ++k) v[k] = s_in[ci][ty + k / 3][tx + k % 3];
#pragma unroll
      for (int k = 0; k < 9; ++k) {
#pragma unroll
        for (int co = 0; co < CO_T; ++co) acc[co] = fmaf(s_w[ci][k][co], v[k], acc[co]);
      }
    }
    __syncthreads();
  }
  const int y = ty0 + ty, x = tx0 + tx;
  if (y < H && x < W) {
#pragma unroll
    for (int co = 0; co < CO_T; ++co) {
      if (co0 + co < Cout) {
        float r = acc[co] + bias[co0 + co];
        if (leaky) r = r > 0.f ? r : 0.2f * r;
        out[((size_t)(b * Cout + co0 + co) * H + y) * W + x] = r;
      }
    }
  }
}

// cat[x, ones * sigma]  (denoiser/base.py:29-30)
__global__ void make_input_simt(const float* __restrict__ x, const float* __restrict__ sigma,
                                int64_t sstride, float* __restrict__ out, int HW, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  int b = i / HW, p = i % HW;
  out[((size_t)b * 2) * HW + p] = x[i];
  out[((size_t)b * 2 + 1) * HW + p] = sigma[b * sstride];
}

__global__ void maxpool2_simt(const float* __restrict__ in, float* __restrict__ out, int BC, int H,
                              int W) {  // nn.MaxPool2d(2), unet.py:83
  int Ho = H / 2, Wo = W / 2;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)BC * Ho * Wo) return;
  int x = i % Wo, y = (i / Wo) % Ho;
  size_t bc = i / ((size_t)Wo * Ho);
  const float* p = in + (bc * H + 2 * y) * W + 2 * x;
  out[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True), unet.py:99
__global__ void upsample2_simt(const float* __restrict__ in, float* __restrict__ out, int BC, int H,
                               int W) {
  int Ho = 2 * H, Wo = 2 * W;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)BC * Ho * Wo) return;
  int x = i % Wo, y = (i / Wo) % Ho;
  size_t bc = i / ((size_t)Wo * Ho);
  float sy = (float)(H - 1) / (float)(Ho - 1), sx = (float)(W - 1) / (float)(Wo - 1);
  float fy = sy * y, fx = sx * x;
  int y0 = (int)fy, x0 = (int)fx;
  int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
  float ly = fy - y0, lx = fx - x0;
  const float* p = in + bc * H * W;
  float top = (1.f - lx) * p[y0 * W + x0] + lx * p[y0 * W + x1];
  float bot = (1.f - lx) * p[y1 * W + x0] + lx * p[y1 * W + x1];
  out[i] = (1.f - ly) * top + ly * bot;
}

// outconv 1x1 (unet.py:124-131) + residual (unet.py:65-66) + clamp (denoiser/base.py:32)
__global__ void outc_simt(const float* __restrict__ feat, const float* __restrict__ w,
                          const float* __restrict__ bias, const float* __restrict__ x,
                          float* __restrict__ out, int C, int HW, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  int b = i / HW, p = i % HW;
  float acc = bias[0];
  for (int c = 0; c < C; ++c) acc = fmaf(w[c], feat[((size_t)b * C + c) * HW + p], acc);
  float r = x[i] + acc;
  out[i] = fminf(fmaxf(r, 0.f), 1.f);
}


// ---- reverse mode: the vector-Jacobian product of the denoiser ---------------------------------
// d/d(input) of a 3x3 pad-1 convolution is the 3x3 pad-1 convolution of the output gradient with the weights
// transposed (Cin <-> Cout) and flipped (tap (ky,kx) -> (2-ky,2-kx)), so conv3x3_simt serves both directions; the
// element-wise adjoints (LeakyReLU, MaxPool2d, bilinear x2, 1x1 outconv + residual + clamp) are the kernels below.

// pre-clamp output r = x + outconv(feat)   (unet.py:65-66, 124-131); the clamp mask needs r, not clamp(r)
__global__ void outc_pre_simt(const float* __restrict__ feat, const float* __restrict__ w, const float* __restrict__ bias,
                              const float* __restrict__ x, float* __restrict__ r, int C, int HW, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  int b = i / HW, p = i % HW;
  float acc = bias[0];
  for (int c = 0; c < C; ++c) acc = fmaf(w[c], feat[((size_t)b * C + c) * HW + p], acc);
  r[i] = x[i] + acc;
}

// g_r = gout * 1[0 <= r <= 1] (torch.clamp's backward);  gpre[b,c,p] = w[c] * g_r * lrelu'(a[b,c,p])
__global__ void outc_bwd_simt(const float* __restrict__ gout, const float* __restrict__ r, const float* __restrict__ w,
                              const float* __restrict__ a, float* __restrict__ gr, float* __restrict__ gpre, int C, int HW,
                              int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * HW) return;
  int b = i / HW, p = i % HW;
  const float rv = r[i];
  const float g = (rv >= 0.f && rv <= 1.f) ? gout[i] : 0.f;
  gr[i] = g;
  for (int c = 0; c < C; ++c) {
    const size_t j = ((size_t)b * C + c) * HW + p;
    gpre[j] = w[c] * g * (a[j] > 0.f ? 1.f : 0.2f);
  }
}

// g *= lrelu'(a): a is the layer's post-activation output (same sign as the pre-activation)
__global__ void lrelu_bwd_simt(float* __restrict__ g, const float* __restrict__ a, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] *= (a[i] > 0.f ? 1.f : 0.2f);
}

// Adjoint of MaxPool2d(2) (first maximum in scan order wins, as in ATen) + the skip-connection gradient + lrelu':
//   gpre[b,c,Y,X] = (argmax(b,c,Y/2,X/2) == (Y,X) ? gpool[b,c,Y/2,X/2] : 0) + gskip[b,c,Y,X]) * lrelu'(a[b,c,Y,X])
// a, gpre: [B,C,H,W]; gpool: [B,C,H/2,W/2]; gskip: channels [0,C) of a [B,Ccat,H,W] tensor (nullable)
__global__ void pool_bwd_simt(const float* __restrict__ gpool, const float* __restrict__ a, const float* __restrict__ gskip,
                              int Ccat, float* __restrict__ gpre, int C, int H, int W, int B) {
  const int Ho = H / 2, Wo = W / 2;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * Ho * Wo) return;
  const int x = i % Wo, y = (i / Wo) % Ho;
  const size_t bc = i / ((size_t)Wo * Ho);
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const size_t base = (bc * H + 2 * y) * W + 2 * x;
  const float v[4] = {a[base], a[base + 1], a[base + W], a[base + W + 1]};
  int arg = 0;
  float best = v[0];
#pragma unroll
  for (int k = 1; k < 4; ++k) if (v[k] > best) { best = v[k]; arg = k; }
  const float gp = gpool[i];
  const size_t sbase = ((b * Ccat + c) * H + 2 * y) * (size_t)W + 2 * x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const size_t off = (size_t)(k >> 1) * W + (k & 1);
    float g = (k == arg ? gp : 0.f) + (gskip ? gskip[sbase + off] : 0.f);
    gpre[base + off] = g * (v[k] > 0.f ? 1.f : 0.2f);
  }
}

// weight of high-resolution index Y on low-resolution index i for bilinear x2 with align_corners=True (the float
// expressions of upsample2_simt, so the adjoint is that of the forward kernel)
__device__ __forceinline__ float up_weight(int Y, int i, int h, float s) {
  const float f = s * Y;
  const int y0 = (int)f;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0);
  const float l = f - y0;
  return (y0 == i ? 1.f - l : 0.f) + (y1 == i ? l : 0.f);
}

// Adjoint of the x2 bilinear up-sampling + lrelu' of the low-resolution source:
//   gpre[b,c,i,j] = lrelu'(a[b,c,i,j]) * sum_{Y,X} wy(Y,i) wx(X,j) gup[b,coff+c,Y,X]
// gup: channels [coff, coff+C) of a [B,Ccat,2h,2w] tensor; a, gpre: [B,C,h,w]
__global__ void up_bwd_simt(const float* __restrict__ gcat, int Ccat, int coff, const float* __restrict__ a,
                            float* __restrict__ gpre, int C, int h, int w, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)B * C * h * w) return;
  const int xj = i % w, yi = (i / w) % h;
  const size_t bc = i / ((size_t)w * h);
  const int c = (int)(bc % C);
  const size_t b = bc / C;
  const int Ho = 2 * h, Wo = 2 * w;
  const float sy = (float)(h - 1) / (float)(Ho - 1), sx = (float)(w - 1) / (float)(Wo - 1);
  const float* g = gcat + (b * Ccat + coff + c) * (size_t)Ho * Wo;
  float wx[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int X = 2 * xj - 2 + k;
    wx[k] = (X >= 0 && X < Wo) ? up_weight(X, xj, w, sx) : 0.f;
  }
  float acc = 0.f;
#pragma unroll
  for (int m = 0; m < 6; ++m) {
    const int Y = 2 * yi - 2 + m;
    if (Y < 0 || Y >= Ho) continue;
    const float wy = up_weight(Y, yi, h, sy);
    if (wy == 0.f) continue;
    float row = 0.f;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int X = 2 * xj - 2 + k;
      if (wx[k] != 0.f) row = fmaf(wx[k], g[(size_t)Y * Wo + X], row);
    }
    acc = fmaf(wy, row, acc);
  }
  gpre[i] = acc * (a[i] > 0.f ? 1.f : 0.2f);
}

// gx = d/d(x) = gin2[:,0] + g_r (the residual connection);  gsigma[b] = sum_p gin2[b,1,p] (the noise map is sigma[b] everywhere)
__global__ void __launch_bounds__(256)
first_bwd_finish_simt(const float* __restrict__ gin2, const float* __restrict__ gr, float* __restrict__ gx,
                      float* __restrict__ gsigma, int64_t gs_stride, int HW) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const float* g0 = gin2 + (size_t)b * 2 * HW;
  const float* g1 = g0 + HW;
  float s = 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) {
    gx[(size_t)b * HW + p] = g0[p] + gr[(size_t)b * HW + p];
    s += g1[p];
  }
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) gsigma[b * gs_stride] = red[0];
}

struct UNetSimt : Denoiser {
  DevBuf weights;   // raw state_dict floats
  size_t w_off[kNumUnetConv3], b_off[kNumUnetConv3], outc_w, outc_b;
  DevBuf ws;        // activation workspace
  size_t ws_B = 0, ws_HW = 0;

  int init(const float* host) {
    TFPNP_TRY(weights.alloc(kUnetParamCount * sizeof(float)));
    TFPNP_CUDA_OK(cudaMemcpy(weights.p, host, kUnetParamCount * sizeof(float), cudaMemcpyHostToDevice));
    host_w.assign(host, host + kUnetParamCount);
    size_t off = 0;
    const ConvSpec* sp = unet_conv_specs();
    for (int l = 0; l < kNumUnetConv3; ++l) {
      w_off[l] = off; off += (size_t)sp[l].cout * sp[l].cin * 9;
      b_off[l] = off; off += sp[l].cout;
    }
    outc_w = off; off += 32;
    outc_b = off; off += 1;
    if (off != kUnetParamCount) { set_error("unet param table mismatch"); return TFPNP_ERR_INVALID; }
    return 0;
  }

  int conv(int l, const float* s0, int C0, const float* s1, int C1, float* out, int B, int H, int W,
           cudaStream_t st) {
    const ConvSpec& sp = unet_conv_specs()[l];
    if (C0 + C1 != sp.cin) { set_error("simt conv %d channel mismatch", l); return TFPNP_ERR_INVALID; }
    dim3 grid(cdiv(W, TS) * cdiv(H, TS), cdiv(sp.cout, CO_T), B);
    const float* wp = weights.as<float>();
    conv3x3_simt<<<grid, TS * TS, 0, st>>>(s0, C0, s1, C1, wp + w_off[l], wp + b_off[l], out, sp.cout, H, W, 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }

  // workspace (floats per image, in units of HW): in 2 | x1 32 | x2 16 | x3 8 | x4 4 | x5 2 |
  // tmpA 64 | tmpB 64 (ping-pong for block-internal activations, pooled and upsampled tensors)
  int prepare(int B, int H, int W) override {
    TFPNP_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "UNet needs H,W multiples of 16, got %dx%d", H, W);
    const size_t per_img = (size_t)(2 + 32 + 16 + 8 + 4 + 2 + 64 + 64) * H * W;
    const void* before = ws.p;
    TFPNP_TRY(ws.alloc(per_img * B * sizeof(float)));
    if (ws.p != before) ++generation;
    return 0;
  }

  int forward(const float* x, const float* sigma, int64_t sstride, float* out, int B, int H, int W,
              cudaStream_t st) override {
    const size_t HW = (size_t)H * W;
    TFPNP_CHECK(ws.bytes >= (size_t)(2 + 32 + 16 + 8 + 4 + 2 + 64 + 64) * HW * B * sizeof(float), "prepare() not called");
    float* base = ws.as<float>();
    float* in2 = base;
    float* x1 = in2 + 2 * HW * B;
    float* x2 = x1 + 32 * HW * B;
    float* x3 = x2 + 16 * HW * B;
    float* x4 = x3 + 8 * HW * B;
    float* x5 = x4 + 4 * HW * B;
    float* tA = x5 + 2 * HW * B;
    float* tB = tA + 64 * HW * B;
    const int T = 256;
    make_input_simt<<<cdiv((int)(B * HW), T), T, 0, st>>>(x, sigma, sstride, in2, (int)HW, B);
    TFPNP_COUNT_LAUNCH();
    // encoder
    TFPNP_TRY(conv(0, in2, 2, nullptr, 0, tA, B, H, W, st));
    TFPNP_TRY(conv(1, tA, 32, nullptr, 0, tB, B, H, W, st));
    TFPNP_TRY(conv(2, tB, 32, nullptr, 0, x1, B, H, W, st));
    float* skips[5] = {x1, x2, x3, x4, x5};
    int ch[5] = {32, 64, 128, 256, 512};
    for (int lv = 1; lv <= 4; ++lv) {
      int h = H >> lv, w = W >> lv;
      size_t n = (size_t)B * ch[lv - 1] * h * w;
      maxpool2_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(skips[lv - 1], tA, B * ch[lv - 1], h * 2, w * 2);
      TFPNP_COUNT_LAUNCH();
      int l0 = 3 * lv;
      TFPNP_TRY(conv(l0, tA, ch[lv - 1], nullptr, 0, tB, B, h, w, st));
      TFPNP_TRY(conv(l0 + 1, tB, ch[lv], nullptr, 0, tA, B, h, w, st));
      TFPNP_TRY(conv(l0 + 2, tA, ch[lv], nullptr, 0, skips[lv], B, h, w, st));
    }
    // decoder
    const float* cur = x5;
    int cur_c = 512;
    for (int k = 0; k < 4; ++k) {
      int lv = 3 - k;  // output level
      int h = H >> lv, w = W >> lv;
      size_t n = (size_t)B * cur_c * h * w;
      upsample2_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(cur, tA, B * cur_c, h / 2, w / 2);
      TFPNP_COUNT_LAUNCH();
      int l0 = 15 + 3 * k;
      float* o0 = tB;
      float* o1 = tA;  // tA (upsampled) is dead after conv-0
      TFPNP_TRY(conv(l0, skips[lv], ch[lv], tA, cur_c, o0, B, h, w, st));
      TFPNP_TRY(conv(l0 + 1, o0, ch[lv], nullptr, 0, o1, B, h, w, st));
      TFPNP_TRY(conv(l0 + 2, o1, ch[lv], nullptr, 0, o0, B, h, w, st));
      cur = o0;
      cur_c = ch[lv];
    }
    const float* wp = weights.as<float>();
    outc_simt<<<cdiv((int)(B * HW), T), T, 0, st>>>(cur, wp + outc_w, wp + outc_b, x, out, 32, (int)HW, B);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }

  // ---- reverse mode ---------------------------------------------------------------------------------------------
  DevBuf wt;        // per layer: weights transposed + flipped, [Cin][Cout][3][3]  (the input-gradient convolution)
  DevBuf zero_bias; // 768 zeros
  DevBuf gws;       // activations of every layer + gradient buffers
  size_t wt_off[kNumUnetConv3];
  std::vector<float> host_w;   // state_dict floats kept for the lazy build of `wt`

  int ensure_grad_weights() {
    if (wt.p) return 0;
    const ConvSpec* sp = unet_conv_specs();
    size_t total = 0;
    for (int l = 0; l < kNumUnetConv3; ++l) { wt_off[l] = total; total += (size_t)sp[l].cout * sp[l].cin * 9; }
    std::vector<float> h(total);
    for (int l = 0; l < kNumUnetConv3; ++l) {
      const int ci = sp[l].cin, co = sp[l].cout;
      const float* w = host_w.data() + w_off[l];
      float* t = h.data() + wt_off[l];
      for (int o = 0; o < co; ++o)
        for (int c = 0; c < ci; ++c)
          for (int k = 0; k < 9; ++k) t[((size_t)c * co + o) * 9 + (8 - k)] = w[((size_t)o * ci + c) * 9 + k];
    }
    TFPNP_TRY(wt.alloc(total * sizeof(float)));
    TFPNP_CUDA_OK(cudaMemcpy(wt.p, h.data(), total * sizeof(float), cudaMemcpyHostToDevice));
    TFPNP_TRY(zero_bias.alloc(768 * sizeof(float)));
    TFPNP_CUDA_OK(cudaMemset(zero_bias.p, 0, 768 * sizeof(float)));
    return 0;
  }

  // input gradient of layer l: gin [B,cout,h,w] -> gout [B,cin,h,w]
  int dgrad(int l, const float* gin, float* gout, int B, int H, int W, cudaStream_t st) {
    const ConvSpec& sp = unet_conv_specs()[l];
    dim3 grid(cdiv(W, TS) * cdiv(H, TS), cdiv(sp.cin, CO_T), B);
    conv3x3_simt<<<grid, TS * TS, 0, st>>>(gin, sp.cout, nullptr, 0, wt.as<float>() + wt_off[l], zero_bias.as<float>(), gout,
                                           sp.cin, H, W, 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }

  // (gx, gsigma) = J(x, sigma)^T gout for out = clamp(x + UNet(cat[x, sigma]), 0, 1): recomputes the forward pass keeping
  // every layer's activation (fp32 NCHW, 366 HW floats per image), then walks the layers backwards.
  int vjp(const float* x, const float* sigma, int64_t sstride, const float* gout, float* gx, float* gsigma,
          int64_t gs_stride, int B, int H, int W, cudaStream_t st) override {
    TFPNP_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "UNet needs H,W multiples of 16, got %dx%d", H, W);
    TFPNP_TRY(ensure_grad_weights());
    const ConvSpec* sp = unet_conv_specs();
    const size_t HW = (size_t)H * W;
    // workspace, in units of HW floats per image
    size_t act_units = 0;
    size_t a_off[kNumUnetConv3];
    for (int l = 0; l < kNumUnetConv3; ++l) { a_off[l] = act_units; act_units += (size_t)sp[l].cout * HW >> (2 * sp[l].level); }
    // (sizes in floats per image from here on)
    const size_t in2_off = act_units, pool_off = in2_off + 2 * HW, up_off = pool_off + 8 * HW, r_off = up_off + 64 * HW,
                 gr_off = r_off + HW, gA_off = gr_off + HW, gB_off = gA_off + 32 * HW, gcat_off = gB_off + 32 * HW;
    const size_t gcat_sz[4] = {96 * HW, 48 * HW, 24 * HW, 12 * HW};   // level 0..3: (skip + up) channels at that level
    const size_t per_img = gcat_off + gcat_sz[0] + gcat_sz[1] + gcat_sz[2] + gcat_sz[3];
    TFPNP_TRY(gws.alloc(per_img * B * sizeof(float)));
    float* base = gws.as<float>();
    float* a[kNumUnetConv3];
    for (int l = 0; l < kNumUnetConv3; ++l) a[l] = base + a_off[l] * B;
    float* in2 = base + in2_off * B;
    float* pooled = base + pool_off * B;
    float* upbuf = base + up_off * B;
    float* r = base + r_off * B;
    float* gr = base + gr_off * B;
    float* gA = base + gA_off * B;
    float* gB = base + gB_off * B;
    float* gcat[4];
    gcat[0] = base + gcat_off * B;
    for (int lv = 1; lv < 4; ++lv) gcat[lv] = gcat[lv - 1] + gcat_sz[lv - 1] * B;
    const int ch[5] = {32, 64, 128, 256, 512};
    const int T = 256;
    const float* wp = weights.as<float>();

    // forward, every activation kept
    make_input_simt<<<cdiv((int)(B * HW), T), T, 0, st>>>(x, sigma, sstride, in2, (int)HW, B);
    TFPNP_COUNT_LAUNCH();
    TFPNP_TRY(conv(0, in2, 2, nullptr, 0, a[0], B, H, W, st));
    TFPNP_TRY(conv(1, a[0], 32, nullptr, 0, a[1], B, H, W, st));
    TFPNP_TRY(conv(2, a[1], 32, nullptr, 0, a[2], B, H, W, st));
    for (int lv = 1; lv <= 4; ++lv) {
      const int h = H >> lv, w = W >> lv, l0 = 3 * lv;
      const size_t n = (size_t)B * ch[lv - 1] * h * w;
      maxpool2_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(a[l0 - 1], pooled, B * ch[lv - 1], h * 2, w * 2);
      TFPNP_COUNT_LAUNCH();
      TFPNP_TRY(conv(l0, pooled, ch[lv - 1], nullptr, 0, a[l0], B, h, w, st));
      TFPNP_TRY(conv(l0 + 1, a[l0], ch[lv], nullptr, 0, a[l0 + 1], B, h, w, st));
      TFPNP_TRY(conv(l0 + 2, a[l0 + 1], ch[lv], nullptr, 0, a[l0 + 2], B, h, w, st));
    }
    for (int k = 0; k < 4; ++k) {
      const int lv = 3 - k, h = H >> lv, w = W >> lv, l0 = 15 + 3 * k;
      const size_t n = (size_t)B * ch[lv + 1] * h * w;
      upsample2_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(a[l0 - 1], upbuf, B * ch[lv + 1], h / 2, w / 2);
      TFPNP_COUNT_LAUNCH();
      TFPNP_TRY(conv(l0, a[3 * lv + 2], ch[lv], upbuf, ch[lv + 1], a[l0], B, h, w, st));
      TFPNP_TRY(conv(l0 + 1, a[l0], ch[lv], nullptr, 0, a[l0 + 1], B, h, w, st));
      TFPNP_TRY(conv(l0 + 2, a[l0 + 1], ch[lv], nullptr, 0, a[l0 + 2], B, h, w, st));
    }
    outc_pre_simt<<<cdiv((int)(B * HW), T), T, 0, st>>>(a[26], wp + outc_w, wp + outc_b, x, r, 32, (int)HW, B);
    TFPNP_COUNT_LAUNCH();

    // backward
    float* cur = gA;      // gradient w.r.t. the pre-activation output of layer l
    float* oth = gB;
    outc_bwd_simt<<<cdiv((int)(B * HW), T), T, 0, st>>>(gout, r, wp + outc_w, a[26], gr, cur, 32, (int)HW, B);
    TFPNP_COUNT_LAUNCH();
    for (int l = 26; l >= 1; --l) {
      const int lv = sp[l].level, h = H >> lv, w = W >> lv;
      if (l >= 15 && (l - 15) % 3 == 0) {
        // decoder block head: input = cat[skip(level lv), up(a[l-1])]  (unet.py:99-121)
        TFPNP_TRY(dgrad(l, cur, gcat[lv], B, h, w, st));
        const size_t n = (size_t)B * ch[lv + 1] * (h / 2) * (w / 2);
        up_bwd_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(gcat[lv], sp[l].cin, ch[lv], a[l - 1], cur, ch[lv + 1], h / 2,
                                                               w / 2, B);
        TFPNP_COUNT_LAUNCH();
      } else if (l <= 12 && l % 3 == 0) {
        // encoder block head: input = maxpool(a[l-1]); a[l-1] is also the skip of level lv-1  (unet.py:80-90)
        TFPNP_TRY(dgrad(l, cur, oth, B, h, w, st));
        const size_t n = (size_t)B * ch[lv - 1] * h * w;
        pool_bwd_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(oth, a[l - 1], gcat[lv - 1], ch[lv - 1] + ch[lv], cur,
                                                                 ch[lv - 1], 2 * h, 2 * w, B);
        TFPNP_COUNT_LAUNCH();
      } else {
        TFPNP_TRY(dgrad(l, cur, oth, B, h, w, st));
        const size_t n = (size_t)B * sp[l].cin * h * w;
        lrelu_bwd_simt<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(oth, a[l - 1], n);
        TFPNP_COUNT_LAUNCH();
        float* t = cur; cur = oth; oth = t;
      }
    }
    TFPNP_TRY(dgrad(0, cur, oth, B, H, W, st));      // [B,2,H,W]: d/dx through the network, d/d(noise map)
    first_bwd_finish_simt<<<B, 256, 0, st>>>(oth, gr, gx, gsigma, gs_stride, (int)HW);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  ~UNetSimt() override { weights.release(); ws.release(); wt.release(); zero_bias.release(); gws.release(); }
};

}  // namespace

Denoiser* make_unet_simt(const float* weights_host) {
  UNetSimt* u = new UNetSimt();
  u->precision = TFPNP_PREC_FP32_SIMT;
  if (u->init(weights_host) != 0) { delete u; return nullptr; }
  return u;
}

}  // namespace tfpnp
