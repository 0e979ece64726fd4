// IRCNN prox_sigma denoiser (BASELINE configs[0] names it).  The reference ships no IRCNN
// (create_denoiser, tfpnp/pnp/__init__.py:5-13, only knows 'unet'; ConvLayer's unused `dilation` argument,
// tfpnp/pnp/denoiser/models/unet.py:9-13, is the only trace), so this is the PUBLISHED network
// (Zhang et al., CVPR 2017, inference form with BatchNorm folded): seven 3x3 convolutions, 64 channels,
// dilations 1,2,3,4,3,2,1, ReLU between them, residual output, wrapped like UNetDenoiser2D
// (tfpnp/pnp/denoiser/base.py:23-32):   out = clamp(x - net(cat[x, sigma * ones]), 0, 1).
// Parity is self-oracled (oracle/pnp_oracle.py: ircnn_denoise) -- "parity unpinned" w.r.t. the reference.
//
//   layer 0   2 -> 64, dil 1, ReLU    CUDA cores (K = 18), NHWC fp16 out
//   layer 1-5 64 -> 64, dil 2,3,4,3,2 tcgen05 implicit GEMM (conv3x3_tc<64> of unet_tc.cu; the dilation is the
//                                     spacing of the nine shifted TMA boxes, zero padding = TMA out-of-bounds fill)
//   layer 6   64 -> 1, dil 1          CUDA cores, fused with the residual and the clamp
#include "common.cuh"
#include <vector>
#include <cstring>

namespace tfpnp {
namespace {

constexpr int kIrcnnDil[7] = {1, 2, 3, 4, 3, 2, 1};

struct IrcnnFirstW { float w[64 * 18]; float b[64]; };

// 3x3 conv over cat[d, sigma*ones] (2 ch) -> 64 ch, ReLU, NHWC fp16 (+ residual plane)
__global__ void __launch_bounds__(128)
ircnn_first_kernel(const float* __restrict__ d, const float* __restrict__ sigma, int64_t sstride,
                   const __grid_constant__ IrcnnFirstW wb, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                   int H, int W) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  const float sg = sigma[b * sstride];
  float in[18];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
    in[k] = ok ? d[((size_t)b * H + yy) * W + xx] : 0.f;
    in[9 + k] = ok ? sg : 0.f;
  }
  const size_t o = (((size_t)b * H + y) * W + x) * 64;
#pragma unroll 1
  for (int c0 = 0; c0 < 64; c0 += 8) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      float a0 = wb.b[c0 + j], a1 = wb.b[c0 + j + 1];
#pragma unroll
      for (int k = 0; k < 18; ++k) {
        a0 = fmaf(wb.w[(c0 + j) * 18 + k], in[k], a0);
        a1 = fmaf(wb.w[(c0 + j + 1) * 18 + k], in[k], a1);
      }
      a0 = fmaxf(a0, 0.f);
      a1 = fmaxf(a1, 0.f);
      __half2 hh = __floats2half2_rn(a0, a1);
      hi[j / 2] = *reinterpret_cast<uint32_t*>(&hh);
      float2 back = __half22float2(hh);
      __half2 ll = __floats2half2_rn((a0 - back.x) * grad_elem::kLoScale, (a1 - back.y) * grad_elem::kLoScale);
      lo[j / 2] = *reinterpret_cast<uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(out_hi + o + c0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (out_lo) *reinterpret_cast<uint4*>(out_lo + o + c0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct IrcnnLastW { float w[9 * 64]; float b; };   // [tap][cin]

// out = clamp(d - (bias + sum_{tap,c} w[tap][c] * act[pixel+tap][c]), 0, 1); one thread per pixel
__global__ void __launch_bounds__(128)
ircnn_last_kernel(const __half* __restrict__ in_hi, const __half* __restrict__ in_lo, const __grid_constant__ IrcnnLastW wb,
                  const float* __restrict__ d, float* __restrict__ out, int H, int W) {
  const int b = blockIdx.z, y = blockIdx.y;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= W) return;
  float acc = wb.b;
#pragma unroll 1
  for (int k = 0; k < 9; ++k) {
    const int yy = y + k / 3 - 1, xx = x + k % 3 - 1;
    if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
    const size_t o = (((size_t)b * H + yy) * W + xx) * 64;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const uint4 v = *reinterpret_cast<const uint4*>(in_hi + o + c8 * 8);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
      float f[8];
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
      if (in_lo) {
        const uint4 vl = *reinterpret_cast<const uint4*>(in_lo + o + c8 * 8);
        const __half2* l = reinterpret_cast<const __half2*>(&vl);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 t = __half22float2(l[i]);
          f[2 * i] = fmaf(t.x, grad_elem::kLoInv, f[2 * i]); f[2 * i + 1] = fmaf(t.y, grad_elem::kLoInv, f[2 * i + 1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(wb.w[k * 64 + c8 * 8 + i], f[i], acc);
    }
  }
  const size_t i = ((size_t)b * H + y) * W + x;
  out[i] = fminf(fmaxf(d[i] - acc, 0.f), 1.f);
}

struct IrcnnTc : Denoiser {
  bool x3 = false;
  IrcnnFirstW first_w;
  IrcnnLastW last_w;
  DevBuf w_hi, w_lo, biases, act;
  size_t w_off[5];
  int pB = 0, pH = 0, pW = 0;
  __half* A[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // ping-pong activations [buffer][plane]
  ConvV1Layer* mid[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};

  // weights_host: model.0.weight [64,2,3,3], model.0.bias [64], model.2.* ... model.10.* [64,64,3,3]+[64],
  // model.12.weight [1,64,3,3], model.12.bias [1]  (flattened in that order)
  int init(const float* host) {
    size_t off = 0;
    for (int i = 0; i < 64 * 18; ++i) first_w.w[i] = host[off + i];
    off += 64 * 18;
    for (int i = 0; i < 64; ++i) first_w.b[i] = host[off + i];
    off += 64;
    const size_t per = (size_t)9 * 64 * 64;
    std::vector<__half> hhi(5 * per), hlo(5 * per);
    std::vector<float> hb(5 * 64);
    for (int l = 0; l < 5; ++l) {
      const float* w = host + off;
      off += (size_t)64 * 64 * 9;
      for (int t = 0; t < 9; ++t)
        for (int o = 0; o < 64; ++o)
          for (int c = 0; c < 64; ++c) {
            const float v = w[((size_t)o * 64 + c) * 9 + t];
            const __half h = __float2half_rn(v);
            const size_t idx = l * per + ((size_t)t * 64 + o) * 64 + c;
            hhi[idx] = h;
            hlo[idx] = __float2half_rn((v - __half2float(h)) * grad_elem::kLoScale);
          }
      for (int i = 0; i < 64; ++i) hb[l * 64 + i] = host[off + i];
      off += 64;
      w_off[l] = l * per;
    }
    for (int c = 0; c < 64; ++c)
      for (int t = 0; t < 9; ++t) last_w.w[t * 64 + c] = host[off + (size_t)c * 9 + t];
    off += 64 * 9;
    last_w.b = host[off];
    off += 1;
    if (off != kIrcnnParamCount) { set_error("ircnn param table mismatch"); return TFPNP_ERR_INVALID; }
    TFPNP_TRY(w_hi.alloc(hhi.size() * sizeof(__half)));
    TFPNP_TRY(biases.alloc(hb.size() * sizeof(float)));
    TFPNP_CUDA_OK(cudaMemcpy(w_hi.p, hhi.data(), hhi.size() * sizeof(__half), cudaMemcpyHostToDevice));
    TFPNP_CUDA_OK(cudaMemcpy(biases.p, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice));
    if (x3) {
      TFPNP_TRY(w_lo.alloc(hlo.size() * sizeof(__half)));
      TFPNP_CUDA_OK(cudaMemcpy(w_lo.p, hlo.data(), hlo.size() * sizeof(__half), cudaMemcpyHostToDevice));
    }
    return 0;
  }

  void free_plans() {
    for (auto& m : mid) { if (m) conv_v1_free(m); m = nullptr; }
  }

  int prepare(int B, int H, int W) override {
    TFPNP_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "IRCNN path needs H,W multiples of 16, got %dx%d", H, W);
    if (B == pB && H == pH && W == pW) return 0;
    free_plans();
    const size_t plane = (size_t)B * H * W * 64;
    const void* before = act.p;
    TFPNP_TRY(act.alloc(plane * 2 * (x3 ? 2 : 1) * sizeof(__half)));
    if (act.p != before) ++generation;
    __half* base = act.as<__half>();
    for (int i = 0; i < 2; ++i) {
      A[i][0] = base + i * plane;
      A[i][1] = x3 ? base + (2 + i) * plane : nullptr;
    }
    for (int l = 0; l < 5; ++l) {
      const int src = l & 1, dst = src ^ 1;
      TFPNP_TRY(conv_v1_plan(&mid[l], A[src][0], A[src][1], 64, w_hi.as<__half>() + w_off[l],
                             x3 ? w_lo.as<__half>() + w_off[l] : nullptr, biases.as<float>() + l * 64, A[dst][0], A[dst][1],
                             B, H, W, 64, kIrcnnDil[l + 1], 0.f));
    }
    pB = B; pH = H; pW = W;
    return 0;
  }

  int forward(const float* x, const float* sigma, int64_t sstride, float* out, int B, int H, int W,
              cudaStream_t st) override {
    TFPNP_CHECK(B == pB && H == pH && W == pW, "prepare(%d,%d,%d) not called (plan is %d,%d,%d)", B, H, W, pB, pH, pW);
    const dim3 grid(cdiv(W, 128), H, B);
    ircnn_first_kernel<<<grid, 128, 0, st>>>(x, sigma, sstride, first_w, A[0][0], A[0][1], H, W);
    TFPNP_COUNT_LAUNCH();
    for (int l = 0; l < 5; ++l) TFPNP_TRY(conv_v1_launch(mid[l], st));
    // five ping-pongs: the last hidden tensor sits in buffer 1
    ircnn_last_kernel<<<grid, 128, 0, st>>>(A[1][0], A[1][1], last_w, x, out, H, W);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }

  ~IrcnnTc() override {
    free_plans();
    w_hi.release(); w_lo.release(); biases.release(); act.release();
  }
};

}  // namespace

Denoiser* make_ircnn_tc(const float* weights_host, int precision) {
  IrcnnTc* n = new IrcnnTc();
  n->precision = precision;
  n->x3 = precision == TFPNP_PREC_FP16X3;
  if (n->init(weights_host) != 0) { delete n; return nullptr; }
  return n;
}

}  // namespace tfpnp
