// Stand-alone entry points: the CT operators and the PSNR reward metric.
#include "tasks.cuh"
#include <map>
#include <memory>

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

struct GeomCache {
  std::map<std::pair<int, int>, std::unique_ptr<CtGeom>> m;
  CtGeom* get(int N, int views, const float* c, const float* s) {
    auto key = std::make_pair(N, views);
    auto it = m.find(key);
    if (it == m.end()) {
      std::unique_ptr<CtGeom> g(new CtGeom());
      if (g->init(N, views) != 0) return nullptr;
      it = m.emplace(key, std::move(g)).first;
    }
    if (c && s && it->second->set_tables(c, s) != 0) return nullptr;
    return it->second.get();
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
  CtGeom* g = geom_cache().get(N, views, cos_host, sin_host);
  if (!g) return TFPNP_ERR_CUDA;
  TFPNP_TRY(g->reserve(B));
  return radon_forward(*g, img, nullptr, sino, B, static_cast<cudaStream_t>(stream));
}

int tfpnp_radon_backward(const float* sino, float* img, int B, int N, int views, const float* cos_host,
                         const float* sin_host, void* stream) {
  TFPNP_CHECK(img && sino && B > 0 && N > 0 && views > 0, "bad argument");
  CtGeom* g = geom_cache().get(N, views, cos_host, sin_host);
  if (!g) return TFPNP_ERR_CUDA;
  return radon_backward(*g, sino, img, B, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
