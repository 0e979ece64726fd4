// Shared host/device helpers for libtfpnp_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include <string>

#include "../../include/tfpnp_b200.h"
#include "grad_elem.cuh"

namespace tfpnp {

// ---- error plumbing (never throw across the C ABI) --------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define TFPNP_CUDA_OK(expr)                                                         \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      ::tfpnp::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,         \
                         cudaGetErrorString(_e));                                   \
      return TFPNP_ERR_CUDA;                                                        \
    }                                                                               \
  } while (0)

#define TFPNP_CHECK(cond, ...)                                                      \
  do {                                                                              \
    if (!(cond)) {                                                                  \
      ::tfpnp::set_error(__VA_ARGS__);                                              \
      return TFPNP_ERR_INVALID;                                                     \
    }                                                                               \
  } while (0)

#define TFPNP_TRY(expr)                                                             \
  do {                                                                              \
    int _s = (expr);                                                                \
    if (_s != 0) return _s;                                                         \
  } while (0)

inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// launch counter: every kernel launch in the library goes through LAUNCH_COUNT()
// so tfpnp_solver_last_launch_count() reports what was really enqueued.
extern thread_local int64_t g_launch_count;
#define TFPNP_COUNT_LAUNCH() (++::tfpnp::g_launch_count)

// Launch with optional programmatic dependent launch (PDL) and thread-block-cluster attributes.
// With pdl = true the kernel may begin (its prologue) before the previous kernel in the stream has
// drained; the kernel itself must execute griddepcontrol.wait before touching dependent data.
template <class... KArgs, class... Args>
inline cudaError_t launch_ex(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                             int cluster, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[2];
  unsigned n = 0;
  if (pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// simple owned device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int alloc(size_t n) {
    if (n <= bytes) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
    cudaError_t e = cudaMalloc(&p, n);
    if (e != cudaSuccess) {
      set_error("cudaMalloc(%zu) failed: %s", n, cudaGetErrorString(e));
      return TFPNP_ERR_NOMEM;
    }
    bytes = n;
    return 0;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Scratch buffer out of a per-(device, stream) pool of cached cudaMalloc blocks (misc.cu).  The reverse-mode entry points
// took ~15 cudaMalloc / cudaFree per call plus a stream synchronisation before freeing; a released block goes back to the
// pool of ITS stream and is only ever handed to later work on that stream, so stream order alone makes the reuse safe and
// no synchronisation is needed.  tfpnp_release_cached_scratch() frees the pools.
void* scratch_pool_take(size_t bytes, cudaStream_t st, size_t* got);
void scratch_pool_give(void* p, size_t bytes, cudaStream_t st);
struct PoolBuf {
  void* p = nullptr;
  size_t bytes = 0;
  cudaStream_t st = nullptr;
  PoolBuf() {}
  PoolBuf(const PoolBuf&) = delete;
  PoolBuf& operator=(const PoolBuf&) = delete;
  ~PoolBuf() { release(); }
  int alloc(size_t n, cudaStream_t stream) {
    release();
    st = stream;
    p = scratch_pool_take(n ? n : 1, st, &bytes);
    if (!p) {
      set_error("scratch allocation of %zu bytes failed", n);
      return TFPNP_ERR_NOMEM;
    }
    return 0;
  }
  void release() {
    if (p) scratch_pool_give(p, bytes, st);
    p = nullptr;
    bytes = 0;
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// kNumUnetConv3, kUnetParamCount, ConvSpec, unet_conv_specs(): grad_elem.cuh (plain C++, shared with the CPU emulation)

// abstract denoiser: d -> clamp(UNet(cat[d, sigma]), 0, 1)
struct Denoiser {
  int precision = 0;
  // bumped whenever prepare() moves the workspaces: CUDA graphs that captured the old
  // addresses must be dropped by their owners
  int generation = 0;
  virtual ~Denoiser() {}
  // allocate workspaces / descriptors for this shape (never called during graph capture)
  virtual int prepare(int B, int H, int W) = 0;
  // a second engine over the SAME weights with its own activation workspace, so two half-batches can run
  // concurrently on two streams (nullptr if the implementation does not support it); owned by the caller
  virtual Denoiser* clone_shared() { return nullptr; }
  // x, out: [B,H,W] fp32 (C=1); sigma[b] at sigma[b*sigma_stride]
  virtual int forward(const float* x, const float* sigma, int64_t sigma_stride, float* out, int B,
                      int H, int W, cudaStream_t st) = 0;
  // reverse mode: gx [B,H,W] = d<out,gout>/dx, gsigma[b*gs_stride] = d<out,gout>/dsigma[b].
  // Implemented by the fp32 engine (unet_simt.cu); the tensor-core engines report TFPNP_ERR_UNSUPPORTED.
  virtual int vjp(const float* x, const float* sigma, int64_t sigma_stride, const float* gout, float* gx,
                  float* gsigma, int64_t gs_stride, int B, int H, int W, cudaStream_t st) {
    (void)x; (void)sigma; (void)sigma_stride; (void)gout; (void)gx; (void)gsigma; (void)gs_stride; (void)B; (void)H; (void)W; (void)st;
    set_error("this denoiser engine has no reverse mode: create it with precision fp32_simt");
    return TFPNP_ERR_UNSUPPORTED;
  }
  // measurement aid: per-launch device times of one forward (tfpnp_denoiser_layer_profile); names: 48 bytes per launch
  virtual int layer_profile(const float* x, const float* sigma, float* out, int B, int H, int W, int reps, float* ms_out,
                            int cap, int* n_out, char* names, cudaStream_t st) {
    (void)x; (void)sigma; (void)out; (void)B; (void)H; (void)W; (void)reps; (void)ms_out; (void)cap; (void)n_out; (void)names; (void)st;
    set_error("this denoiser engine has no per-launch profile");
    return TFPNP_ERR_UNSUPPORTED;
  }
  // debugging aid: the reverse-mode workspace of the last vjp() (device pointer, size in floats); nullptr if none
  virtual const float* grad_workspace(size_t* n_floats) { *n_floats = 0; return nullptr; }
};

// pre-planned one-tile-per-CTA tensor-core 3x3 conv layer (unet_tc.cu): NHWC fp16 (hi [+ lo residual plane]) in/out,
// weights [tap][Cout][Cin] fp16, act = max(v, slope*v), tap spacing `dil`
struct ConvV1Layer;
int conv_v1_plan(ConvV1Layer** out, const __half* x_hi, const __half* x_lo, int Cin, const __half* w_hi,
                 const __half* w_lo, const float* bias, __half* out_hi, __half* out_lo, int B, int H, int W,
                 int Cout, int dil, float slope);
int conv_v1_launch(const ConvV1Layer* L, cudaStream_t st);
void conv_v1_free(ConvV1Layer* L);

Denoiser* make_ircnn_tc(const float* weights_host, int precision);   // ircnn.cu
constexpr size_t kIrcnnParamCount = (size_t)64 * 2 * 9 + 64 + 5 * ((size_t)64 * 64 * 9 + 64) + (size_t)1 * 64 * 9 + 1;

Denoiser* make_unet_simt(const float* weights_host);                 // unet_simt.cu
Denoiser* make_unet_tc(const float* weights_host, int precision);    // unet_tc.cu
int conv3x3_nhwc(const void* x0, int C0, const void* x1, int C1, const void* w_taps, const float* bias,
                 void* out, int B, int H, int W, int Cout, cudaStream_t st);   // unet_tc.cu

}  // namespace tfpnp
