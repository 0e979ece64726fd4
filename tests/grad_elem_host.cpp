// CPU emulation of UNetSimt::vjp (tfpnp_b200/csrc/unet_simt.cu) for tests/test_grad.py.  TEST INFRASTRUCTURE ONLY.
// Runs the SAME layer sequence, workspace layout and per-element adjoint bodies as the CUDA engine
// (grad_elem::unet_vjp_sequence / *_elem from tfpnp_b200/csrc/grad_elem.cuh), with plain loops standing in for the kernel
// launches and a naive convolution standing in for conv3x3_simt.  Built by the test with g++ (no CUDA needed).
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include "../tfpnp_b200/csrc/grad_elem.cuh"

using namespace tfpnp;

namespace {

// out[b,co,y,x] = act(bias[co] + sum w[co,ci,ky,kx] in[b,ci,y+ky-1,x+kx-1]), input = cat(src0, src1); w: [Cout][Cin][9]
void conv3x3_host(const float* s0, int C0, const float* s1, int C1, const float* w, const float* bias, float* out, int Cout,
                  int B, int H, int W, bool leaky) {
  const int Cin = C0 + C1;
  for (int b = 0; b < B; ++b)
    for (int co = 0; co < Cout; ++co) {
      float* o = out + ((size_t)b * Cout + co) * H * W;
      for (int i = 0; i < H * W; ++i) o[i] = bias ? bias[co] : 0.f;
      for (int c = 0; c < Cin; ++c) {
        const float* in = c < C0 ? s0 + ((size_t)b * C0 + c) * H * W : s1 + ((size_t)b * C1 + (c - C0)) * H * W;
        const float* wk = w + ((size_t)co * Cin + c) * 9;
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx) {
            const float wv = wk[ky * 3 + kx];
            const int y_lo = ky == 0 ? 1 : 0, y_hi = ky == 2 ? H - 1 : H;
            const int x_lo = kx == 0 ? 1 : 0, x_hi = kx == 2 ? W - 1 : W;
            for (int y = y_lo; y < y_hi; ++y) {
              const float* ir = in + (size_t)(y + ky - 1) * W + (kx - 1);
              float* orow = o + (size_t)y * W;
              for (int x = x_lo; x < x_hi; ++x) orow[x] += wv * ir[x];
            }
          }
      }
      if (leaky) for (int i = 0; i < H * W; ++i) o[i] = o[i] > 0.f ? o[i] : 0.2f * o[i];
    }
}

// stand-in for the tcgen05 kernel (conv_v1_*): NHWC fp16 in/out, weights [tap][rows][K] fp16, fp32 accumulate,
// out = max(v, slope v) rounded to fp16
// With residual planes (FP16X3): the three products x_hi w_hi + x_lo w_hi + x_hi w_lo, and Y split into hi + residual.
void conv3x3_tc_host(const uint16_t* X, const uint16_t* Xlo, int K, const uint16_t* Wt, const uint16_t* Wlo, const float* bias,
                     float slope, uint16_t* Y, uint16_t* Ylo, int rows, int B, int H, int W) {
  std::vector<float> xf((size_t)B * H * W * K), wf((size_t)9 * rows * K), xl, wl;
  for (size_t i = 0; i < xf.size(); ++i) xf[i] = grad_elem::h2f_bits(X[i]);
  for (size_t i = 0; i < wf.size(); ++i) wf[i] = grad_elem::h2f_bits(Wt[i]);
  if (Xlo) {
    xl.resize(xf.size()); wl.resize(wf.size());
    // residual planes are stored scaled by kLoScale (grad_elem.cuh); the kernels scale the correction sums back
    for (size_t i = 0; i < xf.size(); ++i) xl[i] = grad_elem::h2f_bits(Xlo[i]) * grad_elem::kLoInv;
    for (size_t i = 0; i < wf.size(); ++i) wl[i] = grad_elem::h2f_bits(Wlo[i]) * grad_elem::kLoInv;
  }
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int r = 0; r < rows; ++r) {
          float acc = bias ? bias[r] : 0.f;
          for (int t = 0; t < 9; ++t) {
            const int yy = y + t / 3 - 1, xx = x + t % 3 - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float* xp = xf.data() + (((size_t)b * H + yy) * W + xx) * K;
            const float* wp = wf.data() + ((size_t)t * rows + r) * K;
            float s = 0.f;
            for (int k = 0; k < K; ++k) s += xp[k] * wp[k];
            if (Xlo) {
              const float* xlp = xl.data() + (xp - xf.data());
              const float* wlp = wl.data() + (wp - wf.data());
              for (int k = 0; k < K; ++k) s += xlp[k] * wp[k] + xp[k] * wlp[k];
            }
            acc += s;
          }
          const float v = acc > slope * acc ? acc : slope * acc;
          const size_t o = (((size_t)b * H + y) * W + x) * rows + r;
          Y[o] = grad_elem::f2h_bits(v);
          if (Ylo) Ylo[o] = grad_elem::f2h_bits((v - grad_elem::h2f_bits(Y[o])) * grad_elem::kLoScale);
        }
}

struct HostOps {
  int tc = 0;                              // convolutions through the fp16 (1) / split-fp16 (2) NHWC path; 3: split-fp16 forward + fp16 gradients
  bool bwd = false;                        // the op in flight is an input-gradient convolution
  std::vector<uint16_t> X, Y, Xlo, Ylo;
  bool lo_on() const { return tc == 2 || (tc == 3 && !bwd); }
  uint16_t* lo(std::vector<uint16_t>& v) { return lo_on() ? v.data() : nullptr; }
  std::vector<float> scale;
  const float* flat;                       // state_dict floats
  size_t w_off[kNumUnetConv3], b_off[kNumUnetConv3], outc_w, outc_b;
  std::vector<float> wt;                   // transposed + flipped weights
  size_t wt_off[kNumUnetConv3];
  int B, H, W;

  void init() {
    const ConvSpec* sp = unet_conv_specs();
    size_t off = 0, total = 0;
    for (int l = 0; l < kNumUnetConv3; ++l) {
      w_off[l] = off; off += (size_t)sp[l].cout * sp[l].cin * 9;
      b_off[l] = off; off += sp[l].cout;
      wt_off[l] = total; total += (size_t)sp[l].cout * sp[l].cin * 9;
    }
    outc_w = off; outc_b = off + 32;
    wt.resize(total);
    for (int l = 0; l < kNumUnetConv3; ++l)
      grad_elem::transpose_flip_weights(flat + w_off[l], wt.data() + wt_off[l], sp[l].cout, sp[l].cin);
  }
  int make_input(const float* x, const float* sigma, int64_t sstride, float* in2) {
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b)
      for (size_t p = 0; p < HW; ++p) { in2[(size_t)b * 2 * HW + p] = x[b * HW + p]; in2[((size_t)b * 2 + 1) * HW + p] = sigma[b * sstride]; }
    return 0;
  }
  void to_half(const float* src, int C, int Ctot, int coff, int hw, const float* sc) {
    for (size_t i = 0; i < (size_t)B * C * hw; ++i) grad_elem::to_half_nhwc_elem(i, src, X.data(), lo(Xlo), C, Ctot, coff, hw, sc);
  }
  void from_half(size_t yoff, float* dst, int C, int Ctot, int coff, int hw, const float* sc) {
    for (size_t i = 0; i < (size_t)B * C * hw; ++i)
      grad_elem::from_half_nhwc_elem(i, Y.data() + yoff, lo_on() ? Ylo.data() + yoff : nullptr, dst, C, Ctot, coff, hw, sc);
  }
  void reset_xy() {
    const size_t n = (size_t)96 * H * W * B;
    X.assign(n, 0); Y.assign(n, 0); Xlo.assign(n, 0); Ylo.assign(n, 0);
  }
  int conv(int l, const float* s0, int C0, const float* s1, int C1, float* out, int h, int w) {
    const ConvSpec& sp = unet_conv_specs()[l];
    if (C0 + C1 != sp.cin) return -1;
    if (tc && l >= 1) {                    // UNetSimt::GradOps::conv, tensor-core branch
      bwd = false;
      reset_xy();
      to_half(s0, C0, C0 + C1, 0, h * w, nullptr);
      if (s1) to_half(s1, C1, C0 + C1, C0, h * w, nullptr);
      std::vector<uint16_t> wt16((size_t)9 * sp.cout * sp.cin), wl16(wt16.size());
      grad_elem::build_tc_weights(flat + w_off[l], sp.cout, sp.cin, false, 0, sp.cout, wt16.data(), lo(wl16));
      conv3x3_tc_host(X.data(), lo(Xlo), sp.cin, wt16.data(), lo(wl16), flat + b_off[l], 0.2f, Y.data(), lo(Ylo), sp.cout, B, h, w);
      from_half(0, out, sp.cout, sp.cout, 0, h * w, nullptr);
      return 0;
    }
    conv3x3_host(s0, C0, s1, C1, flat + w_off[l], flat + b_off[l], out, sp.cout, B, h, w, true);
    return 0;
  }
  int maxpool(const float* in, float* out, int C, int h, int w) {
    for (size_t bc = 0; bc < (size_t)B * C; ++bc)
      for (int y = 0; y < h / 2; ++y)
        for (int x = 0; x < w / 2; ++x) {
          const float* p = in + (bc * h + 2 * y) * w + 2 * x;
          out[(bc * (h / 2) + y) * (w / 2) + x] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[w], p[w + 1]));
        }
    return 0;
  }
  int upsample(const float* in, float* out, int C, int h, int w) {      // the expressions of upsample2_simt
    const int Ho = 2 * h, Wo = 2 * w;
    const float sy = (float)(h - 1) / (float)(Ho - 1), sx = (float)(w - 1) / (float)(Wo - 1);
    for (size_t bc = 0; bc < (size_t)B * C; ++bc)
      for (int y = 0; y < Ho; ++y)
        for (int x = 0; x < Wo; ++x) {
          const float fy = sy * y, fx = sx * x;
          const int y0 = (int)fy, x0 = (int)fx;
          const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
          const float ly = fy - y0, lx = fx - x0;
          const float* p = in + bc * h * w;
          const float top = (1.f - lx) * p[y0 * w + x0] + lx * p[y0 * w + x1];
          const float bot = (1.f - lx) * p[y1 * w + x0] + lx * p[y1 * w + x1];
          out[(bc * Ho + y) * Wo + x] = (1.f - ly) * top + ly * bot;
        }
    return 0;
  }
  int outc_pre(const float* a26, const float* x, float* r) {
    const size_t HW = (size_t)H * W;
    for (size_t i = 0; i < (size_t)B * HW; ++i) {
      const size_t b = i / HW, p = i % HW;
      float acc = flat[outc_b];
      for (int c = 0; c < 32; ++c) acc = fmaf(flat[outc_w + c], a26[(b * 32 + c) * HW + p], acc);
      r[i] = x[i] + acc;
    }
    return 0;
  }
  int outc_bwd(const float* gout, const float* r, const float* a26, float* gr, float* gpre) {
    for (size_t i = 0; i < (size_t)B * H * W; ++i) grad_elem::outc_bwd_elem(i, gout, r, flat + outc_w, a26, gr, gpre, 32, H * W);
    return 0;
  }
  int dgrad(int l, const float* gin, float* gout, int h, int w) {
    const ConvSpec& sp = unet_conv_specs()[l];
    if (tc && l >= 1) {                    // UNetSimt::GradOps::dgrad, tensor-core branch
      bwd = true;
      reset_xy();
      scale.assign(B, 1.f);
      const size_t per = (size_t)sp.cout * h * w;
      for (int b = 0; b < B; ++b) {        // absmax_scale_simt
        float m = 0.f;
        for (size_t i = 0; i < per; ++i) m = fmaxf(m, fabsf(gin[b * per + i]));
        scale[b] = grad_elem::pow2_scale(m);
      }
      to_half(gin, sp.cout, sp.cout, 0, h * w, scale.data());
      int rows[2];
      const int np = grad_elem::dgrad_parts(l, rows);
      size_t yoff = 0;
      for (int p = 0, coff = 0, r0 = 0; p < np; coff += rows[p], r0 += rows[p], ++p) {
        std::vector<uint16_t> wt16((size_t)9 * rows[p] * sp.cout), wl16(wt16.size());
        grad_elem::build_tc_weights(flat + w_off[l], sp.cout, sp.cin, true, r0, rows[p], wt16.data(), lo(wl16));
        conv3x3_tc_host(X.data(), lo(Xlo), sp.cout, wt16.data(), lo(wl16), nullptr, 1.0f, Y.data() + yoff,
                        lo_on() ? Ylo.data() + yoff : nullptr, rows[p], B, h, w);
        from_half(yoff, gout, rows[p], sp.cin, coff, h * w, scale.data());
        yoff += (size_t)B * h * w * rows[p];
      }
      return 0;
    }
    conv3x3_host(gin, sp.cout, nullptr, 0, wt.data() + wt_off[l], nullptr, gout, sp.cin, B, h, w, false);
    return 0;
  }
  int lrelu_bwd(float* g, const float* a, size_t n) {
    for (size_t i = 0; i < n; ++i) g[i] *= grad_elem::lrelu_d(a[i]);
    return 0;
  }
  int pool_bwd(const float* gpool, const float* a, const float* gskip, int Ccat, float* gpre, int C, int h, int w) {
    for (size_t i = 0; i < (size_t)B * C * (h / 2) * (w / 2); ++i) grad_elem::pool_bwd_elem(i, gpool, a, gskip, Ccat, gpre, C, h, w);
    return 0;
  }
  int up_bwd(const float* gcat, int Ccat, int coff, const float* a, float* gpre, int C, int h, int w) {
    for (size_t i = 0; i < (size_t)B * C * h * w; ++i) grad_elem::up_bwd_elem(i, gcat, Ccat, coff, a, gpre, C, h, w);
    return 0;
  }
  int first_finish(const float* gin2, const float* gr, float* gx, float* gsigma, int64_t gs_stride) {   // first_bwd_finish_simt
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (size_t p = 0; p < HW; ++p) {
        gx[b * HW + p] = gin2[(size_t)b * 2 * HW + p] + gr[b * HW + p];
        s += gin2[((size_t)b * 2 + 1) * HW + p];
      }
      gsigma[b * gs_stride] = (float)s;
    }
    return 0;
  }
};

// centred ortho 2-D DFT pair of tfpnp/utils/transforms.py:68-103 (fftshift(FFT(ifftshift(x))) for even N) by direct
// summation: X[k] = 1/sqrt(N) sum_n x[n] exp(-+ 2 pi i (k - N/2)(n - N/2) / N) along both axes
void dft2c_host(const cplx* in, cplx* out, int N, bool inverse) {
  std::vector<double> cr((size_t)N * N), ci((size_t)N * N);
  const double sgn = inverse ? 1.0 : -1.0, scale = 1.0 / std::sqrt((double)N);
  for (int k = 0; k < N; ++k)
    for (int n = 0; n < N; ++n) {
      const double ang = sgn * 2.0 * M_PI * (double)((k - N / 2) * (n - N / 2) % N) / N;
      cr[(size_t)k * N + n] = std::cos(ang) * scale; ci[(size_t)k * N + n] = std::sin(ang) * scale;
    }
  std::vector<double> tr((size_t)N * N), ti((size_t)N * N);
  for (int r = 0; r < N; ++r)           // along columns (dim -2 of [H,W,2] is W; order is irrelevant for a separable DFT)
    for (int k = 0; k < N; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double xr = in[(size_t)r * N + n].x, xi = in[(size_t)r * N + n].y;
        ar += xr * cr[(size_t)k * N + n] - xi * ci[(size_t)k * N + n];
        ai += xr * ci[(size_t)k * N + n] + xi * cr[(size_t)k * N + n];
      }
      tr[(size_t)r * N + k] = ar; ti[(size_t)r * N + k] = ai;
    }
  for (int c = 0; c < N; ++c)
    for (int k = 0; k < N; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double xr = tr[(size_t)n * N + c], xi = ti[(size_t)n * N + c];
        ar += xr * cr[(size_t)k * N + n] - xi * ci[(size_t)k * N + n];
        ai += xr * ci[(size_t)k * N + n] + xi * cr[(size_t)k * N + n];
      }
      out[(size_t)k * N + c].x = (float)ar; out[(size_t)k * N + c].y = (float)ai;
    }
}

struct HostAdmmOps {
  const float* flat; const cplx* y0; const uint8_t* mask;   // y0 [B,N,N] complex, mask [B,N,N] in natural (centred) order
  int B, N;
  size_t n() const { return (size_t)B * N * N; }
  int slot_get(const cplx* state, cplx* buf, int k) {
    const size_t HW = (size_t)N * N;
    for (size_t i = 0; i < n(); ++i) buf[i] = state[((i / HW) * 3 + k) * HW + i % HW];
    return 0;
  }
  int slot_put(cplx* state, const cplx* buf, int k) {
    const size_t HW = (size_t)N * N;
    for (size_t i = 0; i < n(); ++i) state[((i / HW) * 3 + k) * HW + i % HW] = buf[i];
    return 0;
  }
  int pre(const cplx* gz, const cplx* gu, const cplx* st_i, const cplx* st_n, cplx* A, cplx* IN) {
    for (size_t i = 0; i < n(); ++i) grad_elem::admm_pre_elem(i, gz, gu, st_i, st_n, A, IN, N * N);
    return 0;
  }
  // the two modes of masked_fft_step (csmri_variants.cu): G = ifft2c(op(fft2c(in)))
  int step(const cplx* in, cplx* out, const float* mu_i, bool blend, bool with_y0) {
    const size_t HW = (size_t)N * N;
    std::vector<cplx> Z(HW);
    for (int b = 0; b < B; ++b) {
      dft2c_host(in + b * HW, Z.data(), N, false);
      for (size_t p = 0; p < HW; ++p) {
        const bool on = mask[b * HW + p] != 0;
        const cplx y = with_y0 ? y0[b * HW + p] : cplx{0.f, 0.f};
        if (blend) {
          if (on) { const float m = mu_i[b]; Z[p].x = (m * Z[p].x + y.x) / (1.f + m); Z[p].y = (m * Z[p].y + y.y) / (1.f + m); }
        } else {
          if (on) { Z[p].x -= y.x; Z[p].y -= y.y; } else { Z[p].x = 0.f; Z[p].y = 0.f; }
        }
      }
      dft2c_host(Z.data(), out + b * HW, N, true);
    }
    return 0;
  }
  int blend(const cplx* A, const float* mu_i, cplx* Q) { return step(A, Q, mu_i, true, false); }
  int residual(const cplx* IN, cplx* R) { return step(IN, R, nullptr, false, true); }
  int mu_reduce(const cplx* A, const cplx* R, const float* mu_i, float* gmu, int64_t stride) {      // grad_mu_reduce
    const size_t HW = (size_t)N * N;
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (size_t p = 0; p < HW; ++p) s += (double)A[b * HW + p].x * R[b * HW + p].x + (double)A[b * HW + p].y * R[b * HW + p].y;
      const float m = 1.f + mu_i[b];
      gmu[b * stride] = (float)s / (m * m);
    }
    return 0;
  }
  int mid(const cplx* gx, cplx* gu, const cplx* Q, const cplx* st_i, float* gxt, float* v) {
    for (size_t i = 0; i < n(); ++i) grad_elem::admm_mid_elem(i, gx, gu, Q, st_i, gxt, v, N * N);
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride);
  int post(const float* gv, cplx* gx, cplx* gz, cplx* gu) {
    for (size_t i = 0; i < n(); ++i) grad_elem::admm_post_elem(i, gv, gx, gz, gu);
    return 0;
  }
};

int unet_vjp_host(const float* weights_flat, const float* x, const float* sigma, int64_t sstride, const float* gout, float* gx,
                  float* gsigma, int64_t gs_stride, int B, int H, int W, int tc = 0, float* ws_out = nullptr) {
  HostOps ops;
  ops.tc = tc;
  ops.flat = weights_flat; ops.B = B; ops.H = H; ops.W = W;
  ops.init();
  std::vector<float> ws(grad_elem::unet_vjp_workspace_floats(B, H, W), 0.f);
  const int rc = grad_elem::unet_vjp_sequence(ops, x, sigma, sstride, gout, gx, gsigma, gs_stride, ws.data(), B, H, W);
  if (ws_out) memcpy(ws_out, ws.data(), ws.size() * sizeof(float));
  return rc;
}

int HostAdmmOps::den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
  return unet_vjp_host(flat, v, sg_i, 1, gxt, gv, gsig, stride, B, N, N);
}

}  // namespace

// states [iters+1][B,3,N,N,2]; y0 [B,1,N,N,2]; mask [B,1,N,N] u8; sigma_d, mu [B,iters] contiguous; outputs as
// tfpnp_csmri_admm_backward (include/tfpnp_b200.h)
extern "C" int emu_admm_backward(const float* weights_flat, const float* states, const float* y0, const uint8_t* mask,
                                 const float* sigma_d, const float* mu, int B, int N, int iters, const float* grad_out,
                                 float* g_sigma, float* g_mu, float* g_state_in) {
  const size_t n = (size_t)B * N * N;
  std::vector<float> P((size_t)2 * B * iters);                       // gather_params3: [sigma | mu][iters][B]
  for (int i = 0; i < iters; ++i)
    for (int b = 0; b < B; ++b) { P[(size_t)i * B + b] = sigma_d[b * iters + i]; P[(size_t)(iters + i) * B + b] = mu[b * iters + i]; }
  std::vector<cplx> c[7];
  for (auto& v : c) v.assign(n, cplx{0.f, 0.f});
  std::vector<float> f[3];
  for (auto& v : f) v.assign(n, 0.f);
  HostAdmmOps ops{weights_flat, reinterpret_cast<const cplx*>(y0), mask, B, N};
  grad_elem::AdmmGradBufs w{c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(), c[5].data(), c[6].data(),
                            f[0].data(), f[1].data(), f[2].data()};
  return grad_elem::admm_backward_sequence(ops, reinterpret_cast<const cplx*>(states), P.data(), B, N * N, iters,
                                           reinterpret_cast<const cplx*>(grad_out), g_sigma, g_mu,
                                           reinterpret_cast<cplx*>(g_state_in), w);
}

extern "C" int emu_unet_vjp(const float* weights_flat, const float* x, const float* sigma, const float* gout, float* gx,
                            float* gsigma, int B, int H, int W) {
  return unet_vjp_host(weights_flat, x, sigma, 1, gout, gx, gsigma, 1, B, H, W);
}

// the same sequence with every convolution through the fp16 NHWC (tensor-core) branch
extern "C" int emu_unet_vjp_tc(const float* weights_flat, const float* x, const float* sigma, const float* gout, float* gx,
                               float* gsigma, int B, int H, int W, int mode) {
  return unet_vjp_host(weights_flat, x, sigma, 1, gout, gx, gsigma, 1, B, H, W, mode);
}

// psnr_bwd_kernel (misc.cu)
extern "C" int emu_psnr_bwd(const float* out, const float* gt, const float* psnr, const float* gpsnr, float* gout, int B, int64_t HW) {
  for (size_t i = 0; i < (size_t)B * HW; ++i) grad_elem::psnr_bwd_elem(i, out, gt, psnr, gpsnr, gout, (size_t)HW);
  return 0;
}

// spi_backward (spi.cu): the shared sequence + element bodies with loops for the launches
namespace {
struct HostSpiOps {
  const float* flat; const float* x0; const float* K; int64_t K_stride; int B, H, W;
  size_t n() const { return (size_t)B * H * W; }
  int slot_get(const float* state, float* buf, int k) {
    const size_t HW = (size_t)H * W;
    for (size_t i = 0; i < n(); ++i) buf[i] = state[((i / HW) * 3 + k) * HW + i % HW];
    return 0;
  }
  int slot_put(float* state, float* buf, int k) {
    const size_t HW = (size_t)H * W;
    for (size_t i = 0; i < n(); ++i) state[((i / HW) * 3 + k) * HW + i % HW] = buf[i];
    return 0;
  }
  int make_v(const float* st_n, float* v) {
    for (size_t i = 0; i < n(); ++i) grad_elem::spi_v_elem(i, st_n, v, H * W);
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gx, float* gv, float* gsig, int64_t stride) {
    return unet_vjp_host(flat, v, sg_i, 1, gx, gv, gsig, stride, B, H, W);
  }
  int step(const float* st_i, const float* mu_i, const float* gv, float* gx, float* gz, float* gu, float* term) {
    for (size_t i = 0; i < n(); ++i) grad_elem::spi_step_elem(i, st_i, x0, K, K_stride, mu_i, gv, gx, gz, gu, term, H * W);
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {      // image_sum_kernel
    const size_t HW = (size_t)H * W;
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (size_t p = 0; p < HW; ++p) s += term[b * HW + p];
      out[b * stride] = (float)s;
    }
    return 0;
  }
};
}  // namespace

extern "C" int emu_spi_backward(const float* weights_flat, const float* states, const float* x0, const float* K, int64_t K_stride,
                                const float* sigma_d, const float* mu, int B, int H, int W, int iters, const float* grad_out,
                                float* g_sigma, float* g_mu, float* g_state_in) {
  const size_t n = (size_t)B * H * W;
  std::vector<float> P((size_t)2 * B * iters);                       // spi_gather_params
  for (int i = 0; i < iters; ++i)
    for (int b = 0; b < B; ++b) { P[(size_t)i * B + b] = sigma_d[b * iters + i]; P[(size_t)(iters + i) * B + b] = mu[b * iters + i]; }
  std::vector<float> f[6];
  for (auto& v : f) v.assign(n, 0.f);
  HostSpiOps ops{weights_flat, x0, K, K_stride, B, H, W};
  grad_elem::SpiGradBufs w{f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data()};
  return grad_elem::spi_backward_sequence(ops, states, P.data(), B, H * W, iters, grad_out, g_sigma, g_mu, g_state_in, w);
}

// ct backward (misc.cu: tfpnp_ct_iadmm_backward): shared sequence + element bodies; the projector pair A^T(A . [- y0]) is a
// callback into the test (the oracle's Radon pair), everything else loops
namespace {
typedef int (*ata_fn)(const float* img, int with_y0, float* out);
struct HostCtOps {
  const float* flat; ata_fn cb; float inv_opnorm2; int B, N;
  int HW() const { return N * N; }
  size_t n() const { return (size_t)B * N * N; }
  int slot_get(const float* state, float* buf, int k) {
    for (size_t i = 0; i < n(); ++i) buf[i] = state[((i / HW()) * 3 + k) * HW() + i % HW()];
    return 0;
  }
  int slot_put(float* state, float* buf, int k) {
    for (size_t i = 0; i < n(); ++i) state[((i / HW()) * 3 + k) * HW() + i % HW()] = buf[i];
    return 0;
  }
  int pre(const float* gz, const float* gu, const float* st_i, float* gzt, float* z) {
    for (size_t i = 0; i < n(); ++i) grad_elem::ct_pre_elem(i, gz, gu, st_i, gzt, z, HW());
    return 0;
  }
  int ata(const float* img, bool with_y0, float* out) { return cb(img, with_y0 ? 1 : 0, out); }
  int mid(const float* st_i, const float* st_n, const float* gzt, const float* w1, const float* w2, const float* mu_i,
          const float* tau_i, float* gx, float* gz, float* gu, float* v, float* t_tau, float* t_mu) {
    for (size_t i = 0; i < n(); ++i)
      grad_elem::ct_mid_elem(i, st_i, st_n, gzt, w1, w2, mu_i, tau_i, inv_opnorm2, gx, gz, gu, v, t_tau, t_mu, HW());
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (int p = 0; p < HW(); ++p) s += term[(size_t)b * HW() + p];
      out[b * stride] = (float)s;
    }
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return unet_vjp_host(flat, v, sg_i, 1, gxt, gv, gsig, stride, B, N, N);
  }
  int post(const float* gv, float* gx, float* gz, float* gu) {
    for (size_t i = 0; i < n(); ++i) grad_elem::ct_post_elem(i, gv, gx, gz, gu);
    return 0;
  }
};
}  // namespace

extern "C" int emu_ct_backward(const float* weights_flat, const float* states, ata_fn cb, float opnorm, const float* sigma_d,
                               const float* mu, const float* tau, int B, int N, int iters, const float* grad_out, float* g_sigma,
                               float* g_mu, float* g_tau, float* g_state_in) {
  const size_t n = (size_t)B * N * N;
  std::vector<float> P((size_t)3 * B * iters);                       // gather_params_t3
  for (int i = 0; i < iters; ++i)
    for (int b = 0; b < B; ++b) {
      P[(size_t)i * B + b] = sigma_d[b * iters + i];
      P[(size_t)(iters + i) * B + b] = mu[b * iters + i];
      P[(size_t)(2 * iters + i) * B + b] = tau[b * iters + i];
    }
  std::vector<float> f[11];
  for (auto& v : f) v.assign(n, 0.f);
  HostCtOps ops{weights_flat, cb, 1.0f / (opnorm * opnorm), B, N};
  grad_elem::CtGradBufs w{f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data(), f[5].data(), f[6].data(), f[7].data(),
                          f[8].data(), f[9].data(), f[10].data()};
  return grad_elem::ct_backward_sequence(ops, states, P.data(), B, N * N, iters, grad_out, g_sigma, g_mu, g_tau, g_state_in, w);
}

// pr backward (pr.cu: tfpnp_pr_iadmm_backward): shared sequence + element bodies; tfpnp_fft2 replaced by a direct DFT
namespace {
// un-centred ortho 2-D DFT of one N x N complex image (torch.fft.fft2 / ifft2 with norm="ortho")
void dft2_plain_host(const cplx* in, cplx* out, int N, bool inverse) {
  std::vector<double> cr((size_t)N * N), ci((size_t)N * N);
  const double sgn = inverse ? 1.0 : -1.0, scale = 1.0 / std::sqrt((double)N);
  for (int k = 0; k < N; ++k)
    for (int n = 0; n < N; ++n) {
      const double ang = sgn * 2.0 * M_PI * (double)((k * n) % N) / N;
      cr[(size_t)k * N + n] = std::cos(ang) * scale; ci[(size_t)k * N + n] = std::sin(ang) * scale;
    }
  std::vector<double> tr((size_t)N * N), ti((size_t)N * N);
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < N; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double xr = in[(size_t)r * N + n].x, xi = in[(size_t)r * N + n].y;
        ar += xr * cr[(size_t)k * N + n] - xi * ci[(size_t)k * N + n];
        ai += xr * ci[(size_t)k * N + n] + xi * cr[(size_t)k * N + n];
      }
      tr[(size_t)r * N + k] = ar; ti[(size_t)r * N + k] = ai;
    }
  for (int c = 0; c < N; ++c)
    for (int k = 0; k < N; ++k) {
      double ar = 0, ai = 0;
      for (int n = 0; n < N; ++n) {
        const double xr = tr[(size_t)n * N + c], xi = ti[(size_t)n * N + c];
        ar += xr * cr[(size_t)k * N + n] - xi * ci[(size_t)k * N + n];
        ai += xr * ci[(size_t)k * N + n] + xi * cr[(size_t)k * N + n];
      }
      out[(size_t)k * N + c].x = (float)ar; out[(size_t)k * N + c].y = (float)ai;
    }
}

struct HostPrOps {
  const float* flat; const cplx* mask; const float* y0; int B, M, N;
  int HW() const { return N * N; }
  size_t n() const { return (size_t)B * N * N; }
  size_t nm() const { return n() * M; }
  int slot_get(const cplx* state, cplx* buf, int k) {
    for (size_t i = 0; i < n(); ++i) buf[i] = state[((i / HW()) * 3 + k) * HW() + i % HW()];
    return 0;
  }
  int slot_put(cplx* state, cplx* buf, int k) {
    for (size_t i = 0; i < n(); ++i) state[((i / HW()) * 3 + k) * HW() + i % HW()] = buf[i];
    return 0;
  }
  int pre(const cplx* gz, const cplx* gu, const cplx* st_i, cplx* gzt, cplx* z) {
    for (size_t i = 0; i < n(); ++i) grad_elem::pr_pre_elem(i, gz, gu, st_i, gzt, z, HW());
    return 0;
  }
  int mul(const cplx* img, cplx* out) {
    for (size_t i = 0; i < nm(); ++i) grad_elem::pr_mul_elem(i, img, mask, out, M, HW());
    return 0;
  }
  int fft(const cplx* in, cplx* out, bool inverse) {
    for (int k = 0; k < B * M; ++k) dft2_plain_host(in + (size_t)k * HW(), out + (size_t)k * HW(), N, inverse);
    return 0;
  }
  int h(cplx* W, cplx* Bc) {
    for (size_t i = 0; i < nm(); ++i) grad_elem::pr_h_elem(i, W, Bc, y0);
    return 0;
  }
  int acc(const cplx* E, cplx* out) {
    for (size_t i = 0; i < n(); ++i) grad_elem::pr_acc_elem(i, E, mask, out, M, HW());
    return 0;
  }
  int mid(const cplx* st_i, const cplx* st_n, const cplx* gzt, const cplx* jc, const cplx* gzv, const float* mu_i,
          const float* tau_i, const cplx* gx, cplx* gz, cplx* gu, float* gxt, float* v, float* t_tau, float* t_mu) {
    for (size_t i = 0; i < n(); ++i)
      grad_elem::pr_mid_elem(i, st_i, st_n, gzt, jc, gzv, mu_i, tau_i, gx, gz, gu, gxt, v, t_tau, t_mu, HW());
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (int p = 0; p < HW(); ++p) s += term[(size_t)b * HW() + p];
      out[b * stride] = (float)s;
    }
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return unet_vjp_host(flat, v, sg_i, 1, gxt, gv, gsig, stride, B, N, N);
  }
  int post(const float* gv, cplx* gx, cplx* gz, cplx* gu) {
    for (size_t i = 0; i < n(); ++i) grad_elem::pr_post_elem(i, gv, gx, gz, gu);
    return 0;
  }
};
}  // namespace

extern "C" int emu_pr_backward(const float* weights_flat, const float* states, const float* y0, const float* mask, int M,
                               const float* sigma_d, const float* mu, const float* tau, int B, int N, int iters,
                               const float* grad_out, float* g_sigma, float* g_mu, float* g_tau, float* g_state_in) {
  const size_t n = (size_t)B * N * N, nm = n * M;
  std::vector<float> P((size_t)3 * B * iters);
  for (int i = 0; i < iters; ++i)
    for (int b = 0; b < B; ++b) {
      P[(size_t)i * B + b] = sigma_d[b * iters + i];
      P[(size_t)(iters + i) * B + b] = mu[b * iters + i];
      P[(size_t)(2 * iters + i) * B + b] = tau[b * iters + i];
    }
  std::vector<cplx> c1[7], cm[4];
  for (auto& v : c1) v.assign(n, cplx{0.f, 0.f});
  for (auto& v : cm) v.assign(nm, cplx{0.f, 0.f});
  std::vector<float> f1[5];
  for (auto& v : f1) v.assign(n, 0.f);
  HostPrOps ops{weights_flat, reinterpret_cast<const cplx*>(mask), y0, B, M, N};
  grad_elem::PrGradBufs w{c1[0].data(), c1[1].data(), c1[2].data(), c1[3].data(), c1[4].data(), c1[5].data(), c1[6].data(),
                          cm[0].data(), cm[1].data(), cm[2].data(), cm[3].data(), f1[0].data(), f1[1].data(), f1[2].data(),
                          f1[3].data(), f1[4].data()};
  return grad_elem::pr_backward_sequence(ops, reinterpret_cast<const cplx*>(states), P.data(), B, N * N, iters,
                                         reinterpret_cast<const cplx*>(grad_out), g_sigma, g_mu, g_tau,
                                         reinterpret_cast<cplx*>(g_state_in), w);
}

// debugging aid (tools/grad_layer_check.py): the emulation's workspace and its region table
extern "C" int emu_unet_vjp_ws(const float* weights_flat, const float* x, const float* sigma, const float* gout, float* gx,
                               float* gsigma, int B, int H, int W, int mode, float* ws_out) {
  return unet_vjp_host(weights_flat, x, sigma, 1, gout, gx, gsigma, 1, B, H, W, mode, ws_out);
}
extern "C" void emu_unet_vjp_layout(int B, int H, int W, size_t* out39) { grad_elem::unet_vjp_workspace_layout(B, H, W, out39); }

// variant backward (csmri_variants.cu: tfpnp_csmri_variant_backward): shared sequence + element bodies
namespace {
struct HostVarOps {
  HostAdmmOps base;                       // masked-FFT steps and the denoiser VJP of the ADMM emulation
  int B, N;
  int HW() const { return N * N; }
  size_t n() const { return (size_t)B * N * N; }
  int slot_get(const cplx* state, cplx* buf, int V, int k) {
    for (size_t i = 0; i < n(); ++i) buf[i] = state[((i / HW()) * V + k) * HW() + i % HW()];
    return 0;
  }
  int slot_put(cplx* state, cplx* buf, int V, int k) {
    for (size_t i = 0; i < n(); ++i) state[((i / HW()) * V + k) * HW() + i % HW()] = buf[i];
    return 0;
  }
  int pre(int algo, const cplx* st_i, const cplx* st_n, const cplx* g1, const cplx* g2, cplx* A, cplx* IN, int V) {
    for (size_t i = 0; i < n(); ++i) grad_elem::var_pre_elem(algo, i, st_i, st_n, g1, g2, A, IN, V, HW());
    return 0;
  }
  int blend0(const cplx* A, const float* mu, cplx* Q) { return base.step(A, Q, mu, true, false); }
  int resid(const cplx* in, bool with_y0, cplx* out) { return base.step(in, out, nullptr, false, with_y0); }
  int mid(int algo, const cplx* st_i, const cplx* st_n, const cplx* A, const cplx* IN, const cplx* Q, const cplx* R, const float* p1,
          const float* p2, cplx* g0, cplx* g1, cplx* g2, float* gxt, float* v, float* t1, float* t2, int V) {
    for (size_t i = 0; i < n(); ++i) grad_elem::var_mid_elem(algo, i, st_i, st_n, A, IN, Q, R, p1, p2, g0, g1, g2, gxt, v, t1, t2, V, HW());
    return 0;
  }
  int den_vjp(const float* v, const float* sg, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return base.den_vjp(v, sg, gxt, gv, gsig, stride);
  }
  int post1(int algo, const float* gv, cplx* A, cplx* g0, cplx* g1) {
    for (size_t i = 0; i < n(); ++i) grad_elem::var_post1_elem(algo, i, gv, A, g0, g1);
    return 0;
  }
  int post2(int algo, const cplx* A, const cplx* Q, const cplx* R, const float* p1, cplx* g0, cplx* g1, float* t1) {
    for (size_t i = 0; i < n(); ++i) grad_elem::var_post2_elem(algo, i, A, Q, R, p1, g0, g1, t1, HW());
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    for (int b = 0; b < B; ++b) {
      double s = 0;
      for (int p = 0; p < HW(); ++p) s += term[(size_t)b * HW() + p];
      out[b * stride] = (float)s;
    }
    return 0;
  }
};
}  // namespace

extern "C" int emu_variant_backward(int algo, const float* weights_flat, const float* states, const float* y0, const uint8_t* mask,
                                    const float* p0, const float* p1, const float* p2, int B, int N, int iters,
                                    const float* grad_out, float* g_p0, float* g_p1, float* g_p2, float* g_state_in) {
  const size_t n = (size_t)B * N * N;
  std::vector<float> P((size_t)3 * B * iters, 0.f);
  const float* ps[3] = {p0, p1, p2};
  for (int k = 0; k < 3; ++k)
    if (ps[k])
      for (int i = 0; i < iters; ++i)
        for (int b = 0; b < B; ++b) P[((size_t)k * iters + i) * B + b] = ps[k][b * iters + i];
  std::vector<cplx> c[7];
  for (auto& v : c) v.assign(n, cplx{0.f, 0.f});
  std::vector<float> f[5];
  for (auto& v : f) v.assign(n, 0.f);
  HostVarOps ops{HostAdmmOps{weights_flat, reinterpret_cast<const cplx*>(y0), mask, B, N}, B, N};
  grad_elem::VarGradBufs w{c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(), c[5].data(), c[6].data(),
                           f[0].data(), f[1].data(), f[2].data(), f[3].data(), f[4].data()};
  return grad_elem::variant_backward_sequence(ops, algo, reinterpret_cast<const cplx*>(states), P.data(), B, N * N, iters,
                                              reinterpret_cast<const cplx*>(grad_out), g_p0, g_p1, g_p2,
                                              reinterpret_cast<cplx*>(g_state_in), w);
}
