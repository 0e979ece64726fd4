// UNet(2,1) denoiser on CUDA cores in fp32 (TFPNP_PREC_FP32_SIMT).
//
// Verification mode: the same network as tfpnp/pnp/denoiser/models/unet.py:34-131
// evaluated with plain FFMA so the tensor-core path (unet_tc.cu) can be checked
// on the GPU against an fp32 result that shares nothing with it but the weights.
// NCHW fp32 activations, direct 3x3 convolution with shared-memory halo tiles.
#include "common.cuh"
#include "grad_elem.cuh"
#include <vector>
#include <cstdlib>

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

// element bodies: grad_elem.cuh (host+device, exercised on the CPU by tests/test_grad.py)
__global__ void outc_bwd_simt(const float* __restrict__ gout, const float* __restrict__ r, const float* __restrict__ w,
                              const float* __restrict__ a, float* __restrict__ gr, float* __restrict__ gpre, int C, int HW,
                              int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)B * HW) grad_elem::outc_bwd_elem(i, gout, r, w, a, gr, gpre, C, HW);
}

// g *= lrelu'(a): a is the layer's post-activation output (same sign as the pre-activation)
__global__ void lrelu_bwd_simt(float* __restrict__ g, const float* __restrict__ a, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) g[i] *= grad_elem::lrelu_d(a[i]);
}

// adjoint of MaxPool2d(2) + skip-connection gradient + lrelu'; one thread per pooled element
__global__ void pool_bwd_simt(const float* __restrict__ gpool, const float* __restrict__ a, const float* __restrict__ gskip,
                              int Ccat, float* __restrict__ gpre, int C, int H, int W, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)B * C * (H / 2) * (W / 2)) grad_elem::pool_bwd_elem(i, gpool, a, gskip, Ccat, gpre, C, H, W);
}

// adjoint of the x2 bilinear up-sampling + lrelu' of its low-resolution source; one thread per low-resolution element
__global__ void up_bwd_simt(const float* __restrict__ gcat, int Ccat, int coff, const float* __restrict__ a,
                            float* __restrict__ gpre, int C, int h, int w, int B) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)B * C * h * w) grad_elem::up_bwd_elem(i, gcat, Ccat, coff, a, gpre, C, h, w);
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

// ---- tensor-core convolutions inside the reverse-mode sequences (TFPNP_GRAD_TC=1; grad_elem.cuh) -------------------------
// scale[b] = power of two that puts max|g[b]| into [512, 1024); one CTA per image
__global__ void __launch_bounds__(256)
absmax_scale_simt(const float* __restrict__ g, float* __restrict__ scale, size_t per_image) {
  __shared__ float red[256];
  const float* p = g + (size_t)blockIdx.x * per_image;
  float m = 0.f;
  for (size_t i = threadIdx.x; i < per_image; i += 256) m = fmaxf(m, fabsf(p[i]));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + k]);
    __syncthreads();
  }
  if (threadIdx.x == 0) scale[blockIdx.x] = grad_elem::pow2_scale(red[0]);
}
__global__ void to_half_nhwc_simt(const float* __restrict__ src, uint16_t* __restrict__ dst, uint16_t* __restrict__ dst_lo, int C,
                                  int Ctot, int coff, int HW, const float* __restrict__ scale, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::to_half_nhwc_elem(i, src, dst, dst_lo, C, Ctot, coff, HW, scale);
}
__global__ void from_half_nhwc_simt(const uint16_t* __restrict__ src, const uint16_t* __restrict__ src_lo, float* __restrict__ dst,
                                    int C, int Ctot, int coff, int HW, const float* __restrict__ scale, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::from_half_nhwc_elem(i, src, src_lo, dst, C, Ctot, coff, HW, scale);
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


  // tensor-core convolutions for the reverse-mode sequences (opt-in: TFPNP_GRAD_TC=1)
  int grad_tc = 0;              // 0: fp32 CUDA cores; 1: tcgen05 fp16; 2: tcgen05 split-fp16 (FP16X3: hi + residual planes);
                                // 3: split-fp16 forward recompute (it decides the LeakyReLU / max-pool / clamp switches, which is what
                                //    the gradient is sensitive to) + plain fp16 gradient convolutions -- the recommended mix
  int tc_mode_built = 0;
  DevBuf tc_w, tc_wlo, tc_x, tc_xlo, tc_y, tc_ylo, tc_scale;
  size_t tcw_f[kNumUnetConv3], tcw_b[kNumUnetConv3][2];
  ConvV1Layer* tc_fwd[kNumUnetConv3] = {};
  ConvV1Layer* tc_bwd[kNumUnetConv3][2] = {};
  int tcB = 0, tcH = 0, tcW = 0;

  void free_tc_plans() {
    for (int l = 0; l < kNumUnetConv3; ++l) {
      if (tc_fwd[l]) conv_v1_free(tc_fwd[l]);
      tc_fwd[l] = nullptr;
      for (int p = 0; p < 2; ++p) { if (tc_bwd[l][p]) conv_v1_free(tc_bwd[l][p]); tc_bwd[l][p] = nullptr; }
    }
    tcB = tcH = tcW = 0;
  }

  int ensure_tc(int B, int H, int W) {
    const ConvSpec* sp = unet_conv_specs();
    const bool x3 = grad_tc >= 2;            // residual planes exist (forward recompute: modes 2, 3)
    const bool x3f = grad_tc >= 2;           // forward recompute in split-fp16
    const bool x3b = grad_tc == 2;           // gradient convolutions in split-fp16
    if (tc_mode_built != grad_tc) { free_tc_plans(); tc_w.release(); tc_wlo.release(); tc_mode_built = grad_tc; }
    if (!tc_w.p) {
      size_t total = 0;
      for (int l = 1; l < kNumUnetConv3; ++l) {
        tcw_f[l] = total; total += (size_t)9 * sp[l].cout * sp[l].cin;
        int rows[2];
        const int np = grad_elem::dgrad_parts(l, rows);
        for (int p = 0; p < np; ++p) { tcw_b[l][p] = total; total += (size_t)9 * rows[p] * sp[l].cout; }
      }
      std::vector<uint16_t> h(total), hl(x3 ? total : 0);
      for (int l = 1; l < kNumUnetConv3; ++l) {
        const float* w = host_w.data() + w_off[l];
        grad_elem::build_tc_weights(w, sp[l].cout, sp[l].cin, false, 0, sp[l].cout, h.data() + tcw_f[l],
                                    x3 ? hl.data() + tcw_f[l] : nullptr);
        int rows[2];
        const int np = grad_elem::dgrad_parts(l, rows);
        for (int p = 0, r0 = 0; p < np; r0 += rows[p], ++p)
          grad_elem::build_tc_weights(w, sp[l].cout, sp[l].cin, true, r0, rows[p], h.data() + tcw_b[l][p],
                                      x3 ? hl.data() + tcw_b[l][p] : nullptr);
      }
      TFPNP_TRY(tc_w.alloc(total * sizeof(uint16_t)));
      TFPNP_CUDA_OK(cudaMemcpy(tc_w.p, h.data(), total * sizeof(uint16_t), cudaMemcpyHostToDevice));
      if (x3) {
        TFPNP_TRY(tc_wlo.alloc(total * sizeof(uint16_t)));
        TFPNP_CUDA_OK(cudaMemcpy(tc_wlo.p, hl.data(), total * sizeof(uint16_t), cudaMemcpyHostToDevice));
      }
    }
    if (B == tcB && H == tcH && W == tcW) return 0;
    free_tc_plans();
    const size_t HW = (size_t)H * W;
    TFPNP_TRY(tc_x.alloc(96 * HW * B * sizeof(uint16_t)));     // widest operand: cat[skip 32, up 64] at full resolution
    TFPNP_TRY(tc_y.alloc(96 * HW * B * sizeof(uint16_t)));     // widest result: its input gradient, in two parts
    TFPNP_TRY(tc_scale.alloc(B * sizeof(float)));
    if (x3) {
      TFPNP_TRY(tc_xlo.alloc(96 * HW * B * sizeof(uint16_t)));
      TFPNP_TRY(tc_ylo.alloc(96 * HW * B * sizeof(uint16_t)));
    }
    const __half* W16 = tc_w.as<__half>();
    const __half* W16lo = x3 ? tc_wlo.as<__half>() : nullptr;
    const __half* Xlo = x3 ? tc_xlo.as<__half>() : nullptr;
    __half* Ylo = x3 ? tc_ylo.as<__half>() : nullptr;
    const float* biases = weights.as<float>();
    for (int l = 1; l < kNumUnetConv3; ++l) {
      const int h = H >> sp[l].level, w = W >> sp[l].level;
      TFPNP_TRY(conv_v1_plan(&tc_fwd[l], tc_x.as<__half>(), x3f ? Xlo : nullptr, sp[l].cin, W16 + tcw_f[l],
                             x3f ? W16lo + tcw_f[l] : nullptr, biases + b_off[l], tc_y.as<__half>(), x3f ? Ylo : nullptr, B, h, w,
                             sp[l].cout, 1, 0.2f));
      int rows[2];
      const int np = grad_elem::dgrad_parts(l, rows);
      size_t yoff = 0;
      for (int p = 0; p < np; ++p) {
        TFPNP_TRY(conv_v1_plan(&tc_bwd[l][p], tc_x.as<__half>(), x3b ? Xlo : nullptr, sp[l].cout, W16 + tcw_b[l][p],
                               x3b ? W16lo + tcw_b[l][p] : nullptr, zero_bias.as<float>(), tc_y.as<__half>() + yoff,
                               x3b ? Ylo + yoff : nullptr, B, h, w, rows[p], 1, 1.0f));
        yoff += (size_t)B * h * w * rows[p];
      }
    }
    tcB = B; tcH = H; tcW = W;
    return 0;
  }

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
      grad_elem::transpose_flip_weights(w, t, co, ci);
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
  // every layer's activation (fp32 NCHW, 366 HW floats per image), then walks the layers backwards.  The layer sequence
  // and the workspace layout are grad_elem::unet_vjp_sequence (shared with the CPU emulation of tests/test_grad.py).
  struct GradOps {
    UNetSimt* u; int B, H, W; cudaStream_t st;
    static constexpr int T = 256;
    unsigned blocks(size_t n) const { return (unsigned)((n + T - 1) / T); }
    int make_input(const float* x, const float* sigma, int64_t sstride, float* in2) {
      make_input_simt<<<blocks((size_t)B * H * W), T, 0, st>>>(x, sigma, sstride, in2, H * W, B);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    bool lo_planes(bool backward) const { return u->grad_tc == 2 || (u->grad_tc == 3 && !backward); }
    int to_half(const float* src, int C, int Ctot, int coff, int hw, const float* scale, bool backward) {
      const size_t n = (size_t)B * C * hw;
      to_half_nhwc_simt<<<blocks(n), T, 0, st>>>(src, u->tc_x.as<uint16_t>(), lo_planes(backward) ? u->tc_xlo.as<uint16_t>() : nullptr,
                                                 C, Ctot, coff, hw, scale, n);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int from_half(size_t yoff, float* dst, int C, int Ctot, int coff, int hw, const float* scale, bool backward) {
      const size_t n = (size_t)B * C * hw;
      from_half_nhwc_simt<<<blocks(n), T, 0, st>>>(u->tc_y.as<uint16_t>() + yoff,
                                                   lo_planes(backward) ? u->tc_ylo.as<uint16_t>() + yoff : nullptr, dst, C, Ctot, coff,
                                                   hw, scale, n);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int conv(int l, const float* s0, int C0, const float* s1, int C1, float* out, int h, int w) {
      if (!u->grad_tc || l == 0) return u->conv(l, s0, C0, s1, C1, out, B, h, w, st);
      // tcgen05 path: cat[s0, s1] -> NHWC fp16 -> conv + bias + LeakyReLU -> fp32 NCHW
      TFPNP_TRY(to_half(s0, C0, C0 + C1, 0, h * w, nullptr, false));
      if (s1) TFPNP_TRY(to_half(s1, C1, C0 + C1, C0, h * w, nullptr, false));
      TFPNP_TRY(conv_v1_launch(u->tc_fwd[l], st));
      const int cout = unet_conv_specs()[l].cout;
      return from_half(0, out, cout, cout, 0, h * w, nullptr, false);
    }
    int maxpool(const float* in, float* out, int C, int h, int w) {
      maxpool2_simt<<<blocks((size_t)B * C * (h / 2) * (w / 2)), T, 0, st>>>(in, out, B * C, h, w);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int upsample(const float* in, float* out, int C, int h, int w) {
      upsample2_simt<<<blocks((size_t)B * C * 4 * h * w), T, 0, st>>>(in, out, B * C, h, w);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int outc_pre(const float* a26, const float* x, float* r) {
      const float* wp = u->weights.as<float>();
      outc_pre_simt<<<blocks((size_t)B * H * W), T, 0, st>>>(a26, wp + u->outc_w, wp + u->outc_b, x, r, 32, H * W, B);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int outc_bwd(const float* gout, const float* r, const float* a26, float* gr, float* gpre) {
      outc_bwd_simt<<<blocks((size_t)B * H * W), T, 0, st>>>(gout, r, u->weights.as<float>() + u->outc_w, a26, gr, gpre, 32,
                                                              H * W, B);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int dgrad(int l, const float* gin, float* gout, int h, int w) {
      if (!u->grad_tc || l == 0) return u->dgrad(l, gin, gout, B, h, w, st);
      // tcgen05 path: per-image power-of-two scale -> NHWC fp16 -> conv with transposed, flipped weights (one launch per
      // output-channel part) -> fp32 NCHW channel ranges, un-scaled
      const ConvSpec& sp = unet_conv_specs()[l];
      float* scale = u->tc_scale.as<float>();
      absmax_scale_simt<<<B, 256, 0, st>>>(gin, scale, (size_t)sp.cout * h * w);
      TFPNP_COUNT_LAUNCH();
      TFPNP_TRY(to_half(gin, sp.cout, sp.cout, 0, h * w, scale, true));
      int rows[2];
      const int np = grad_elem::dgrad_parts(l, rows);
      size_t yoff = 0;
      for (int p = 0, coff = 0; p < np; coff += rows[p], ++p) {
        TFPNP_TRY(conv_v1_launch(u->tc_bwd[l][p], st));
        TFPNP_TRY(from_half(yoff, gout, rows[p], sp.cin, coff, h * w, scale, true));
        yoff += (size_t)B * h * w * rows[p];
      }
      return 0;
    }
    int lrelu_bwd(float* g, const float* a, size_t n) {
      lrelu_bwd_simt<<<blocks(n), T, 0, st>>>(g, a, n);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int pool_bwd(const float* gpool, const float* a, const float* gskip, int Ccat, float* gpre, int C, int h, int w) {
      pool_bwd_simt<<<blocks((size_t)B * C * (h / 2) * (w / 2)), T, 0, st>>>(gpool, a, gskip, Ccat, gpre, C, h, w, B);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int up_bwd(const float* gcat, int Ccat, int coff, const float* a, float* gpre, int C, int h, int w) {
      up_bwd_simt<<<blocks((size_t)B * C * h * w), T, 0, st>>>(gcat, Ccat, coff, a, gpre, C, h, w, B);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
    int first_finish(const float* gin2, const float* gr, float* gx, float* gsigma, int64_t gs_stride) {
      first_bwd_finish_simt<<<B, 256, 0, st>>>(gin2, gr, gx, gsigma, gs_stride, H * W);
      TFPNP_COUNT_LAUNCH();
      return 0;
    }
  };

  size_t gws_floats = 0;
  const float* grad_workspace(size_t* n_floats) override { *n_floats = gws_floats; return gws.as<float>(); }

  int vjp(const float* x, const float* sigma, int64_t sstride, const float* gout, float* gx, float* gsigma,
          int64_t gs_stride, int B, int H, int W, cudaStream_t st) override {
    TFPNP_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "UNet needs H,W multiples of 16, got %dx%d", H, W);
    TFPNP_TRY(ensure_grad_weights());
    gws_floats = grad_elem::unet_vjp_workspace_floats(B, H, W);
    TFPNP_TRY(gws.alloc(gws_floats * sizeof(float)));
    {
      const char* e = getenv("TFPNP_GRAD_TC");          // tensor-core convolutions: opt-in until validated on a GPU
      grad_tc = e ? atoi(e) : 0;                        // 1: fp16 operands, 2: split-fp16 (FP16X3), 3: split-fp16 forward + fp16 gradients
      if (grad_tc < 0 || grad_tc > 3) grad_tc = 0;
    }
    if (grad_tc) TFPNP_TRY(ensure_tc(B, H, W));
    GradOps ops{this, B, H, W, st};
    TFPNP_TRY(grad_elem::unet_vjp_sequence(ops, x, sigma, sstride, gout, gx, gsigma, gs_stride, gws.as<float>(), B, H, W));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }
  ~UNetSimt() override {
    free_tc_plans();
    for (DevBuf* b : {&weights, &ws, &wt, &zero_bias, &gws, &tc_w, &tc_wlo, &tc_x, &tc_xlo, &tc_y, &tc_ylo, &tc_scale}) b->release();
  }
};

}  // namespace

Denoiser* make_unet_simt(const float* weights_host) {
  UNetSimt* u = new UNetSimt();
  u->precision = TFPNP_PREC_FP32_SIMT;
  if (u->init(weights_host) != 0) { delete u; return nullptr; }
  return u;
}

}  // namespace tfpnp
