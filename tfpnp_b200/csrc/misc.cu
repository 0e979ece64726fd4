// Stand-alone entry points: the CT operators and the PSNR reward metric.
#include <utility>
#include "tasks.cuh"
#include <map>
#include <memory>
#include <mutex>
#include <tuple>
#include <vector>
#include <algorithm>

namespace tfpnp {
namespace {

// psnr[b] = 10 log10(1 / mean((clamp(out,0,1) - gt)^2))   (tfpnp/env/base.py:237-242)
// one CTA per image; fp32 pairwise-ish reduction (warp shuffles + smem)
__global__ void __launch_bounds__(256)
psnr_kernel(const float* __restrict__ out, const float* __restrict__ gt, float* __restrict__ psnr, int64_t HW) {
  const int b = blockIdx.x;
  const float* o = out + (size_t)b * HW;
  const float* g = gt + (size_t)b * HW;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < HW; i += blockDim.x) {
    float v = fminf(fmaxf(o[i], 0.f), 1.f) - g[i];
    acc = fmaf(v, v, acc);
  }
  __shared__ float part[8];
  for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w];
    float mse = t / (float)HW;
    psnr[b] = 10.f * log10f(1.f / mse);
  }
}

// d psnr / d out (grad_elem.cuh: psnr_bwd_elem)
__global__ void psnr_bwd_kernel(const float* __restrict__ out, const float* __restrict__ gt, const float* __restrict__ psnr,
                                const float* __restrict__ gpsnr, float* __restrict__ gout, size_t HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::psnr_bwd_elem(i, out, gt, psnr, gpsnr, gout, HW);
}

// ---- reverse mode of the CT solver: sequence and element bodies in grad_elem.cuh --------------------------
__global__ void real_slot_copy(const float* __restrict__ state, float* __restrict__ buf, int k, int HW, size_t n, int to_state) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float* s = const_cast<float*>(state) + ((i / HW) * 3 + k) * HW + i % HW;
  if (to_state) *s = buf[i]; else buf[i] = *s;
}
__global__ void ct_pre_kernel(const float* __restrict__ GZ, const float* __restrict__ GU, const float* __restrict__ st_i,
                              float* __restrict__ GZT, float* __restrict__ Z, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::ct_pre_elem(i, GZ, GU, st_i, GZT, Z, HW);
}
__global__ void ct_mid_kernel(const float* __restrict__ st_i, const float* __restrict__ st_n, const float* __restrict__ GZT,
                              const float* __restrict__ W1, const float* __restrict__ W2, const float* __restrict__ mu,
                              const float* __restrict__ tau, float inv_opnorm2, float* __restrict__ GX, float* __restrict__ GZ,
                              float* __restrict__ GU, float* __restrict__ v, float* __restrict__ t_tau,
                              float* __restrict__ t_mu, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::ct_mid_elem(i, st_i, st_n, GZT, W1, W2, mu, tau, inv_opnorm2, GX, GZ, GU, v, t_tau, t_mu, HW);
}
__global__ void ct_post_kernel(const float* __restrict__ gv, float* __restrict__ GX, float* __restrict__ GZ,
                               float* __restrict__ GU, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::ct_post_elem(i, gv, GX, GZ, GU);
}
__global__ void __launch_bounds__(256)
image_sum(const float* __restrict__ term, float* __restrict__ out, int64_t stride, int HW) {
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
__global__ void gather_params_t3(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                                 int64_t rs, int64_t cs, float* __restrict__ P, int B, int iters) {
  const int n = B * iters;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = t / B, b = t % B;
    const int64_t src = b * rs + i * cs;
    P[t] = p0[src]; P[n + t] = p1[src]; P[2 * n + t] = p2[src];
  }
}

struct CtGradOps {
  Denoiser* den; const CtGeom* g; const float* y0; float* sino; float inv_opnorm2; int B; cudaStream_t st;
  static constexpr int T = 256;
  int HW() const { return g->N * g->N; }
  size_t n() const { return (size_t)B * HW(); }
  unsigned nb() const { return (unsigned)((n() + T - 1) / T); }
  int slot_get(const float* state, float* buf, int k) {
    real_slot_copy<<<nb(), T, 0, st>>>(state, buf, k, HW(), n(), 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int slot_put(float* state, float* buf, int k) {
    real_slot_copy<<<nb(), T, 0, st>>>(state, buf, k, HW(), n(), 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int pre(const float* gz, const float* gu, const float* st_i, float* gzt, float* z) {
    ct_pre_kernel<<<nb(), T, 0, st>>>(gz, gu, st_i, gzt, z, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int ata(const float* img, bool with_y0, float* out) {      // out = A^T (A img [- y0])
    TFPNP_TRY(radon_forward(*g, img, with_y0 ? y0 : nullptr, sino, B, st));
    return radon_backward(*g, sino, out, B, st);
  }
  int mid(const float* st_i, const float* st_n, const float* gzt, const float* w1, const float* w2, const float* mu_i,
          const float* tau_i, float* gx, float* gz, float* gu, float* v, float* t_tau, float* t_mu) {
    ct_mid_kernel<<<nb(), T, 0, st>>>(st_i, st_n, gzt, w1, w2, mu_i, tau_i, inv_opnorm2, gx, gz, gu, v, t_tau, t_mu, HW(), n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    image_sum<<<B, 256, 0, st>>>(term, out, stride, HW());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return den->vjp(v, sg_i, 1, gxt, gv, gsig, stride, B, g->N, g->N, st);
  }
  int post(const float* gv, float* gx, float* gz, float* gu) {
    ct_post_kernel<<<nb(), T, 0, st>>>(gv, gx, gz, gu, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
};

// Geometry handles behind the stand-alone Radon entry points (tfpnp_radon_forward / _backward, tfpnp_ct_iadmm_backward).
// One entry per (device, N, views, stream): the tables and the transpose scratch live on the device that asked, and two
// streams never share a scratch buffer.  The tables are uploaded when the entry is created or when the caller's tables
// differ from the cached host copy -- not on every call.  Guarded by a mutex (DataParallel-style use is thread-per-GPU).
struct GeomCache {
  struct Entry {
    CtGeom g;
    std::vector<float> c, s;     // host copy of the uploaded tables (empty = the default linspace tables)
  };
  std::mutex mu;
  std::map<std::tuple<int, int, int, cudaStream_t>, std::unique_ptr<Entry>> m;
  CtGeom* get(int N, int views, const float* c, const float* s, cudaStream_t st, int reserve_B) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("cudaGetDevice failed"); return nullptr; }
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(dev, N, views, st);
    auto it = m.find(key);
    if (it == m.end()) {
      std::unique_ptr<Entry> e(new Entry());
      if (e->g.init(N, views) != 0) return nullptr;
      it = m.emplace(key, std::move(e)).first;
    }
    Entry& e = *it->second;
    if (c && s) {
      const bool same = e.c.size() == (size_t)views && std::equal(c, c + views, e.c.begin()) &&
                        std::equal(s, s + views, e.s.begin());
      if (!same) {
        // work queued on this stream may still read the old tables
        if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("stream sync before a table upload failed"); return nullptr; }
        if (e.g.set_tables(c, s) != 0) return nullptr;
        e.c.assign(c, c + views);
        e.s.assign(s, s + views);
      }
    }
    if (reserve_B > 0 && e.g.tbuf.bytes < e.g.scratch_bytes(reserve_B)) {
      if (cudaStreamSynchronize(st) != cudaSuccess) { set_error("stream sync before growing the scratch failed"); return nullptr; }
      if (e.g.reserve(reserve_B) != 0) return nullptr;
    }
    return &e.g;
  }
};
GeomCache& geom_cache() { static GeomCache c; return c; }

}  // namespace
}  // namespace tfpnp

using namespace tfpnp;

extern "C" {

int tfpnp_psnr(const float* out, const float* gt, float* psnr, int B, int64_t HW, void* stream) {
  TFPNP_CHECK(out && gt && psnr && B > 0 && HW > 0, "bad argument");
  psnr_kernel<<<B, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, gt, psnr, HW);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int tfpnp_psnr_backward(const float* out, const float* gt, const float* psnr, const float* grad_psnr, float* grad_out, int B,
                        int64_t HW, void* stream) {
  TFPNP_CHECK(out && gt && psnr && grad_psnr && grad_out && B > 0 && HW > 0, "bad argument");
  const size_t n = (size_t)B * HW;
  psnr_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, gt, psnr, grad_psnr, grad_out,
                                                                                            (size_t)HW, n);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int tfpnp_conv3x3_nhwc(const void* x0, int C0, const void* x1, int C1, const void* w_taps, const float* bias,
                       void* out, int B, int H, int W, int Cout, void* stream) {
  TFPNP_CHECK(x0 && w_taps && bias && out && B > 0, "bad argument");
  return conv3x3_nhwc(x0, C0, x1, C1, w_taps, bias, out, B, H, W, Cout, static_cast<cudaStream_t>(stream));
}

int tfpnp_radon_forward(const float* img, float* sino, int B, int N, int views, const float* cos_host,
                        const float* sin_host, void* stream) {
  TFPNP_CHECK(img && sino && B > 0 && N > 0 && views > 0, "bad argument");
  CtGeom* g = geom_cache().get(N, views, cos_host, sin_host, static_cast<cudaStream_t>(stream), B);
  if (!g) return TFPNP_ERR_CUDA;
  return radon_forward(*g, img, nullptr, sino, B, static_cast<cudaStream_t>(stream));
}

int tfpnp_radon_backward(const float* sino, float* img, int B, int N, int views, const float* cos_host,
                         const float* sin_host, void* stream) {
  TFPNP_CHECK(img && sino && B > 0 && N > 0 && views > 0, "bad argument");
  CtGeom* g = geom_cache().get(N, views, cos_host, sin_host, static_cast<cudaStream_t>(stream), 0);
  if (!g) return TFPNP_ERR_CUDA;
  return radon_backward(*g, sino, img, B, static_cast<cudaStream_t>(stream));
}

int tfpnp_ct_iadmm_backward(void* denoiser, const float* states, const float* y0, int views, float opnorm,
                            const float* cos_host, const float* sin_host, const float* sigma_d, const float* mu,
                            const float* tau, int64_t row_stride, int64_t col_stride, int B, int N, int iters,
                            const float* grad_out, float* grad_sigma_d, float* grad_mu, float* grad_tau, float* grad_state_in,
                            void* stream) {
  TFPNP_CHECK(denoiser && states && y0 && sigma_d && mu && tau && grad_out && grad_sigma_d && grad_mu && grad_tau && B > 0 &&
                  iters > 0 && views > 0 && opnorm > 0.f, "bad argument");
  g_launch_count = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CtGeom* g = geom_cache().get(N, views, cos_host, sin_host, st, B);
  if (!g) return TFPNP_ERR_CUDA;
  const size_t n = (size_t)B * N * N;
  PoolBuf bufs[11], sino, P;
  auto body = [&]() -> int {
    for (PoolBuf& b : bufs) TFPNP_TRY(b.alloc(n * sizeof(float), st));
    TFPNP_TRY(sino.alloc((size_t)B * g->views * g->det * sizeof(float), st));
    TFPNP_TRY(P.alloc((size_t)B * iters * 3 * sizeof(float), st));
    gather_params_t3<<<cdiv(B * iters, 256), 256, 0, st>>>(sigma_d, mu, tau, row_stride, col_stride, P.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    CtGradOps ops{static_cast<Denoiser*>(denoiser), g, y0, sino.as<float>(), 1.0f / (opnorm * opnorm), B, st};
    float* f[11];
    for (int k = 0; k < 11; ++k) f[k] = bufs[k].as<float>();
    grad_elem::CtGradBufs w{f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10]};
    TFPNP_TRY(grad_elem::ct_backward_sequence(ops, states, P.as<float>(), B, N * N, iters, grad_out, grad_sigma_d, grad_mu,
                                              grad_tau, grad_state_in, w));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  };
  const int rc = body();
  for (PoolBuf& b : bufs) b.release();
  sino.release(); P.release();
  return rc;
}

}  // extern "C"

// ---- scratch pool (common.cuh: PoolBuf) ------------------------------------------------------------------------------------
namespace tfpnp {
namespace {
struct ScratchPool {
  std::mutex mu;
  std::map<std::pair<int, cudaStream_t>, std::multimap<size_t, void*>> free_blocks;
  void release_all() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& kv : free_blocks) {
      cudaSetDevice(kv.first.first);
      if (!kv.second.empty()) cudaStreamSynchronize(kv.first.second);   // (a destroyed stream reports an error: ignored)
      for (auto& b : kv.second) cudaFree(b.second);
    }
    cudaGetLastError();
    free_blocks.clear();
  }
};
ScratchPool& scratch_pool() { static ScratchPool* p = new ScratchPool; return *p; }   // leaked on purpose: no teardown order issues
}  // namespace

void* scratch_pool_take(size_t bytes, cudaStream_t st, size_t* got) {
  int dev = 0;
  cudaGetDevice(&dev);
  const size_t want = (bytes + 511) & ~(size_t)511;
  ScratchPool& P = scratch_pool();
  {
    std::lock_guard<std::mutex> lk(P.mu);
    auto& m = P.free_blocks[{dev, st}];
    auto it = m.lower_bound(want);
    if (it != m.end() && it->first <= 2 * want + (1u << 20)) {   // best fit, but never park a huge block under a small request
      void* p = it->second;
      *got = it->first;
      m.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    P.release_all();                                             // cached blocks may be what is in the way
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  *got = want;
  return p;
}

void scratch_pool_give(void* p, size_t bytes, cudaStream_t st) {
  int dev = 0;
  cudaGetDevice(&dev);
  ScratchPool& P = scratch_pool();
  std::lock_guard<std::mutex> lk(P.mu);
  P.free_blocks[{dev, st}].emplace(bytes, p);
}
}  // namespace tfpnp

extern "C" int tfpnp_release_cached_scratch(void) {
  tfpnp::scratch_pool().release_all();
  return 0;
}
