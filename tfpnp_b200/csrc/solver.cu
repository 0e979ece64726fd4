// Solver handle: owns the resident state / workspaces and drives
//   for i in range(iters): x = D(Re(z-u), sigma_i); (z,u) = data-fidelity + dual update
// (tasks/*/solver.py forward loops) as a fixed kernel sequence, optionally replayed from a
// CUDA graph keyed by (B, iters) so one `solver(inputs, parameters)` call is ONE graph launch
// plus an unpack and a pack kernel that touch the caller's tensors.
#include "tasks.cuh"
#include <map>
#include <cstdlib>
#include <vector>

namespace tfpnp {

thread_local int64_t g_launch_count = 0;
static thread_local std::string g_err;
void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}
const char* get_error() { return g_err.c_str(); }

namespace {

// ---- unpack / pack between the reference's state layout and the resident buffers ----
// complex state: [B,3,HW,2] ; real state: [B,3,HW]      (tfpnp/pnp/solver/base.py:95-104)
__global__ void unpack_complex(const float2* __restrict__ st, float* __restrict__ x, float2* __restrict__ z,
                               float2* __restrict__ u, float* __restrict__ d, int HW) {
  int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float2* s = st + (size_t)b * 3 * HW;
    float2 xx = s[i], zz = s[HW + i], uu = s[2 * HW + i];
    size_t o = (size_t)b * HW + i;
    x[o] = xx.x; z[o] = zz; u[o] = uu; d[o] = zz.x - uu.x;
  }
}
__global__ void unpack_real(const float* __restrict__ st, float* __restrict__ x, float* __restrict__ z,
                            float* __restrict__ u, float* __restrict__ d, int HW) {
  int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    const float* s = st + (size_t)b * 3 * HW;
    float xx = s[i], zz = s[HW + i], uu = s[2 * HW + i];
    size_t o = (size_t)b * HW + i;
    x[o] = xx; z[o] = zz; u[o] = uu; d[o] = zz - uu;
  }
}
__global__ void pack_complex(float2* __restrict__ st, const float* __restrict__ x, const float2* __restrict__ z,
                             const float2* __restrict__ u, int HW) {
  int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float2* s = st + (size_t)b * 3 * HW;
    size_t o = (size_t)b * HW + i;
    s[i] = make_float2(x[o], 0.f);   // real2complex (transforms.py:12-13)
    s[HW + i] = z[o];
    s[2 * HW + i] = u[o];
  }
}
__global__ void pack_real(float* __restrict__ st, const float* __restrict__ x, const float* __restrict__ z,
                          const float* __restrict__ u, int HW) {
  int b = blockIdx.y;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
    float* s = st + (size_t)b * 3 * HW;
    size_t o = (size_t)b * HW + i;
    s[i] = x[o]; s[HW + i] = z[o]; s[2 * HW + i] = u[o];
  }
}
// params[b, i] (strided) -> P[k][i][b], k in {sigma_d, mu, tau}; also K*10 for SPI
__global__ void gather_params(const float* __restrict__ sig, const float* __restrict__ mu,
                              const float* __restrict__ tau, int64_t rs, int64_t cs, float* __restrict__ P,
                              int B, int iters, const float* __restrict__ K, int64_t kstride,
                              float* __restrict__ K10) {
  int n = B * iters;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    int i = t / B, b = t % B;
    int64_t src = b * rs + i * cs;
    P[t] = sig[src];
    P[n + t] = mu[src];
    if (tau) P[2 * n + t] = tau[src];
  }
  if (K) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x)
      K10[b] = K[b * kstride] * 10.0f;   // K = K[:,0,0,0] * 10 (tasks/spi/solver.py:32)
  }
}

struct Solver {
  tfpnp_solver_config cfg{};
  Denoiser* den = nullptr;
  Denoiser* den2 = nullptr;          // second engine over the same weights: the batch is denoised as two
                                     // independent halves on two streams, so one half's kernels fill the SMs the
                                     // other half's kernel tails / under-filled deep layers leave idle
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int split_min_batch = 16;
  bool complex_state = false;
  int cap_B = 0, cap_it = 0;
  DevBuf x, z, u, d, T, aux0p, aux1p, params, k10, resid;
  CtGeom geom;
  std::map<std::pair<int, int>, cudaGraphExec_t> graphs;
  std::map<std::pair<int, int>, cudaGraphExec_t> graphs_prof;   // the same loop captured WITH event-record nodes (profiling == 2)
  cudaStream_t cap_stream = nullptr;
  int64_t last_launches = 0;
  int profiling = 0;       // 0 off; 1 eager launches with events; 2 events recorded by nodes of the captured graph
  std::vector<cudaEvent_t> events;
  int prof_iters = 0;
  int den_generation = -1;

  int ensure(int B, int iters) {
    const size_t HW = (size_t)cfg.H * cfg.W;
    const size_t el = complex_state ? sizeof(float2) : sizeof(float);
    if (B > cap_B) {
      // growing the workspaces invalidates captured graphs (they bake the addresses)
      drop_graphs();
      TFPNP_TRY(x.alloc(B * HW * sizeof(float)));
      TFPNP_TRY(d.alloc(B * HW * sizeof(float)));
      TFPNP_TRY(z.alloc(B * HW * el));
      TFPNP_TRY(u.alloc(B * HW * el));
      TFPNP_TRY(k10.alloc(B * sizeof(float)));
      switch (cfg.task) {
        case TFPNP_TASK_CSMRI:
          TFPNP_TRY(T.alloc(B * HW * sizeof(float2)));
          TFPNP_TRY(aux0p.alloc(B * HW * sizeof(float2)));
          TFPNP_TRY(aux1p.alloc(B * HW));
          break;
        case TFPNP_TASK_PR:
          TFPNP_TRY(T.alloc(B * HW * cfg.n_masks * sizeof(float2)));
          TFPNP_TRY(aux0p.alloc(B * HW * cfg.n_masks * sizeof(float)));
          TFPNP_TRY(aux1p.alloc(B * HW * cfg.n_masks * sizeof(float2)));
          break;
        case TFPNP_TASK_CT:
          TFPNP_TRY(resid.alloc((size_t)B * geom.views * geom.det * sizeof(float)));
          TFPNP_TRY(geom.reserve(B));
          TFPNP_TRY(aux0p.alloc((size_t)B * geom.views * geom.det * sizeof(float)));
          break;
        case TFPNP_TASK_SPI:
          TFPNP_TRY(aux0p.alloc(B * HW * sizeof(float)));
          break;
      }
      cap_B = B;
    }
    const int need_it = iters > cap_it ? iters : cap_it;
    if ((size_t)cap_B * need_it * 3 * sizeof(float) > params.bytes) {
      drop_graphs();
      TFPNP_TRY(params.alloc((size_t)cap_B * need_it * 3 * sizeof(float)));
    }
    cap_it = need_it;
    return 0;
  }

  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    for (auto& g : graphs_prof) cudaGraphExecDestroy(g.second);
    graphs.clear();
    graphs_prof.clear();
  }

  bool split_batch(int B) const { return den2 != nullptr && B >= split_min_batch; }

  int denoise(const float* sig, int B, cudaStream_t st) {
    const size_t HW = (size_t)cfg.H * cfg.W;
    if (!split_batch(B)) return den->forward(d.as<float>(), sig, 1, x.as<float>(), B, cfg.H, cfg.W, st);
    const int B0 = B / 2, B1 = B - B0;
    TFPNP_CUDA_OK(cudaEventRecord(ev_fork, st));
    TFPNP_CUDA_OK(cudaStreamWaitEvent(side, ev_fork, 0));
    TFPNP_TRY(den->forward(d.as<float>(), sig, 1, x.as<float>(), B0, cfg.H, cfg.W, st));
    TFPNP_TRY(den2->forward(d.as<float>() + B0 * HW, sig + B0, 1, x.as<float>() + B0 * HW, B1, cfg.H, cfg.W, side));
    TFPNP_CUDA_OK(cudaEventRecord(ev_join, side));
    TFPNP_CUDA_OK(cudaStreamWaitEvent(st, ev_join, 0));
    return 0;
  }

  // enqueue the iteration loop on `st` (no allocation, no sync: graph-capturable)
  // prof_external: the loop is being captured -- record the events as event-record NODES of the graph
  // (cudaEventRecordExternal); a plain cudaEventRecord on a capturing stream only creates a capture-internal dependency
  bool prof_external = false;
  void rec(cudaEvent_t e, cudaStream_t st) {
    if (prof_external) cudaEventRecordWithFlags(e, st, cudaEventRecordExternal);
    else cudaEventRecord(e, st);
  }
  int enqueue_loop(int B, int iters, cudaStream_t st, bool prof) {
    const int N = cfg.W;
    const size_t HW = (size_t)cfg.H * cfg.W;
    const size_t n = (size_t)B * iters;
    const float* P = params.as<float>();
    int ev = 0;
    for (int i = 0; i < iters; ++i) {
      const float* sig = P + (size_t)i * B;
      const float* mu = P + n + (size_t)i * B;
      const float* tau = P + 2 * n + (size_t)i * B;
      if (prof) rec(events[ev++], st);
      if (cfg.task == TFPNP_TASK_SPI) {
        TFPNP_TRY(spi_update(x.as<float>(), z.as<float>(), u.as<float>(), d.as<float>(),
                             aux0p.as<float>(), k10.as<float>(), mu, B, (int)HW, st));
        if (prof) rec(events[ev++], st);
        TFPNP_TRY(denoise(sig, B, st));
        if (prof) rec(events[ev++], st);
        continue;
      }
      TFPNP_TRY(denoise(sig, B, st));
      if (prof) rec(events[ev++], st);
      switch (cfg.task) {
        case TFPNP_TASK_CSMRI:
          TFPNP_TRY(csmri_update(x.as<float>(), z.as<float2>(), u.as<float2>(), d.as<float>(), T.as<float2>(),
                                 aux0p.as<float2>(), aux1p.as<uint8_t>(), mu, B, N, st));
          break;
        case TFPNP_TASK_PR:
          TFPNP_TRY(pr_update(x.as<float>(), z.as<float2>(), u.as<float2>(), d.as<float>(), T.as<float2>(),
                              aux0p.as<float>(), aux1p.as<float2>(), mu, tau, B, cfg.n_masks, N, st));
          break;
        case TFPNP_TASK_CT:
          TFPNP_TRY(ct_update(geom, x.as<float>(), z.as<float>(), u.as<float>(), d.as<float>(), resid.as<float>(),
                              aux0p.as<float>(), 1.0f / (cfg.opnorm * cfg.opnorm), mu, tau, B, st));
          break;
      }
      if (prof) rec(events[ev++], st);
    }
    return 0;
  }

  int forward(const float* state_in, const void* aux0, const void* aux1, int64_t aux1_stride,
              const float* sigma_d, const float* mu, const float* tau, int64_t rs, int64_t cs, int B,
              int iters, float* state_out, cudaStream_t st) {
    TFPNP_CHECK(B > 0 && iters >= 0, "bad B=%d iters=%d", B, iters);
    TFPNP_CHECK(state_in && state_out && state_in != state_out, "state pointers invalid/aliased");
    const bool needs_tau = cfg.task == TFPNP_TASK_PR || cfg.task == TFPNP_TASK_CT;
    TFPNP_CHECK(iters == 0 || (sigma_d && mu && (!needs_tau || tau)), "missing hyper-parameter pointer");
    TFPNP_CHECK(aux0 != nullptr && (cfg.task == TFPNP_TASK_CT || aux1 != nullptr), "missing aux input");
    const int HW = cfg.H * cfg.W;
    g_launch_count = 0;
    TFPNP_TRY(ensure(B, iters > 0 ? iters : 1));
    if (split_batch(B)) {
      TFPNP_TRY(den->prepare(B / 2, cfg.H, cfg.W));
      TFPNP_TRY(den2->prepare(B - B / 2, cfg.H, cfg.W));
    } else {
      TFPNP_TRY(den->prepare(B, cfg.H, cfg.W));
    }
    const int gen = den->generation + (den2 ? den2->generation : 0);
    if (gen != den_generation) {      // a denoiser workspace moved
      drop_graphs();
      den_generation = gen;
    }
    const int T256 = 256;
    dim3 grid(cdiv(HW, T256 * 4), B);
    // 1. bring the caller's tensors into the resident layout
    if (complex_state)
      unpack_complex<<<grid, T256, 0, st>>>(reinterpret_cast<const float2*>(state_in), x.as<float>(),
                                            z.as<float2>(), u.as<float2>(), d.as<float>(), HW);
    else
      unpack_real<<<grid, T256, 0, st>>>(state_in, x.as<float>(), z.as<float>(), u.as<float>(), d.as<float>(), HW);
    TFPNP_COUNT_LAUNCH();
    if (iters > 0) {
      const float* K = cfg.task == TFPNP_TASK_SPI ? static_cast<const float*>(aux1) : nullptr;
      gather_params<<<cdiv(B * iters, T256), T256, 0, st>>>(sigma_d, mu, needs_tau ? tau : nullptr, rs, cs,
                                                             params.as<float>(), B, iters, K, aux1_stride,
                                                             k10.as<float>());
      TFPNP_COUNT_LAUNCH();
      if (cfg.task == TFPNP_TASK_CSMRI)
        TFPNP_TRY(csmri_prep(static_cast<const float*>(aux0), static_cast<const uint8_t*>(aux1),
                             aux0p.as<float2>(), aux1p.as<uint8_t>(), B, cfg.W, st));
      // aux inputs are staged into resident buffers so captured graphs never see the caller's
      // (per-call, allocator-owned) pointers
      if (cfg.task == TFPNP_TASK_PR) {
        TFPNP_TRY(pr_prep(static_cast<const float*>(aux0), aux0p.as<float>(), B, cfg.n_masks, cfg.W, st));
        TFPNP_CUDA_OK(cudaMemcpyAsync(aux1p.p, aux1, (size_t)B * HW * cfg.n_masks * sizeof(float2),
                                      cudaMemcpyDeviceToDevice, st));
      }
      if (cfg.task == TFPNP_TASK_CT)
        TFPNP_CUDA_OK(cudaMemcpyAsync(aux0p.p, aux0, (size_t)B * geom.views * geom.det * sizeof(float),
                                      cudaMemcpyDeviceToDevice, st));
      if (cfg.task == TFPNP_TASK_SPI)
        TFPNP_CUDA_OK(cudaMemcpyAsync(aux0p.p, aux0, (size_t)B * HW * sizeof(float), cudaMemcpyDeviceToDevice, st));
    }
    // 2. the iteration loop
    const bool graphable = cfg.use_graph && profiling != 1 && iters > 0;
    if (graphable) {
      const bool gp = profiling == 2;
      auto& graphs = gp ? this->graphs_prof : this->graphs;
      if (gp) {
        size_t need = (size_t)iters * 3 + 2;
        while (events.size() < need) { cudaEvent_t e; TFPNP_CUDA_OK(cudaEventCreate(&e)); events.push_back(e); }
        prof_iters = iters;
      }
      auto key = std::make_pair(B, iters);
      auto it = graphs.find(key);
      if (it == graphs.end()) {
        if (!cap_stream) TFPNP_CUDA_OK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        int64_t before = g_launch_count;
        TFPNP_CUDA_OK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        prof_external = gp;
        int rc = enqueue_loop(B, iters, cap_stream, gp);
        prof_external = false;
        cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
        if (rc != 0) { if (graph) cudaGraphDestroy(graph); return rc; }
        TFPNP_CUDA_OK(ce);
        cudaGraphExec_t exec = nullptr;
        TFPNP_CUDA_OK(cudaGraphInstantiate(&exec, graph, 0));
        cudaGraphDestroy(graph);
        graphs[key] = exec;
        graph_nodes[key] = g_launch_count - before;
        g_launch_count = before;
        it = graphs.find(key);
      }
      TFPNP_CUDA_OK(cudaGraphLaunch(it->second, st));
      g_launch_count += graph_nodes[key];
    } else {
      if (profiling == 1) {
        size_t need = (size_t)iters * 3 + 2;
        while (events.size() < need) { cudaEvent_t e; TFPNP_CUDA_OK(cudaEventCreate(&e)); events.push_back(e); }
        prof_iters = iters;
      }
      TFPNP_TRY(enqueue_loop(B, iters, st, profiling == 1));
    }
    // 3. hand the state back in the reference layout
    if (complex_state)
      pack_complex<<<grid, T256, 0, st>>>(reinterpret_cast<float2*>(state_out), x.as<float>(), z.as<float2>(),
                                          u.as<float2>(), HW);
    else
      pack_real<<<grid, T256, 0, st>>>(state_out, x.as<float>(), z.as<float>(), u.as<float>(), HW);
    TFPNP_COUNT_LAUNCH();
    TFPNP_CUDA_OK(cudaGetLastError());
    last_launches = g_launch_count;
    return 0;
  }

  int get_profile(float* den_ms, float* upd_ms) {
    *den_ms = 0; *upd_ms = 0;
    if (!profiling || prof_iters == 0) return 0;
    TFPNP_CUDA_OK(cudaEventSynchronize(events[prof_iters * 3 - 1]));
    const bool spi = cfg.task == TFPNP_TASK_SPI;
    for (int i = 0; i < prof_iters; ++i) {
      float a = 0, b = 0;
      TFPNP_CUDA_OK(cudaEventElapsedTime(&a, events[3 * i], events[3 * i + 1]));
      TFPNP_CUDA_OK(cudaEventElapsedTime(&b, events[3 * i + 1], events[3 * i + 2]));
      if (spi) { *upd_ms += a; *den_ms += b; } else { *den_ms += a; *upd_ms += b; }
    }
    return 0;
  }

  std::map<std::pair<int, int>, int64_t> graph_nodes;

  ~Solver() {
    delete den2;
    if (side) cudaStreamDestroy(side);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    drop_graphs();
    for (auto e : events) cudaEventDestroy(e);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    x.release(); z.release(); u.release(); d.release(); T.release(); aux0p.release(); aux1p.release();
    params.release(); k10.release(); resid.release(); geom.cs.release(); geom.sn.release(); geom.tbuf.release();
  }
};

}  // namespace

int solver_create(const tfpnp_solver_config* cfg, Denoiser* den, void** out) {
  TFPNP_CHECK(cfg && den && out, "null argument");
  TFPNP_CHECK(cfg->task >= 0 && cfg->task <= 3, "unknown task %d", cfg->task);   // NotImplementedError upstream
  TFPNP_CHECK(cfg->H == cfg->W, "square images only (got %dx%d)", cfg->H, cfg->W);
  TFPNP_CHECK(cfg->H % 16 == 0 && cfg->H >= 16, "H,W must be multiples of 16");
  const bool fft_task = cfg->task == TFPNP_TASK_CSMRI || cfg->task == TFPNP_TASK_PR;
  TFPNP_CHECK(!fft_task || cfg->H == 32 || cfg->H == 64 || cfg->H == 128 || cfg->H == 256,
              "FFT tasks support N in {32,64,128,256}, got %d", cfg->H);
  TFPNP_CHECK(cfg->task != TFPNP_TASK_PR || (cfg->n_masks >= 1 && cfg->n_masks <= 8), "PR needs 1..8 masks");
  Solver* s = new Solver();
  s->cfg = *cfg;
  s->den = den;
  s->complex_state = fft_task;
  {
    const char* e = getenv("TFPNP_SPLIT");
    const int split = e ? atoi(e) : 0;   // measured: no gain on B200 (persistent grids serialise), off by default
    if (split) {
      s->den2 = den->clone_shared();
      if (s->den2) {
        if (cudaStreamCreateWithFlags(&s->side, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_join, cudaEventDisableTiming) != cudaSuccess) {
          delete s; set_error("stream/event creation failed"); return TFPNP_ERR_CUDA;
        }
      }
    }
  }
  if (cfg->task == TFPNP_TASK_CT) {
    if (!(cfg->views > 0 && cfg->opnorm > 0.f)) { delete s; set_error("CT needs views > 0 and opnorm > 0"); return TFPNP_ERR_INVALID; }
    int rc = s->geom.init(cfg->H, cfg->views);
    if (rc == 0 && cfg->ct_cos && cfg->ct_sin) rc = s->geom.set_tables(cfg->ct_cos, cfg->ct_sin);
    if (rc != 0) { delete s; return rc; }
  }
  *out = s;
  return 0;
}

}  // namespace tfpnp

// =============================== C ABI =========================================
using namespace tfpnp;

extern "C" {

int tfpnp_version(void) { return TFPNP_B200_VERSION; }
const char* tfpnp_last_error(void) { return get_error(); }

int tfpnp_denoiser_create(const float* weights_host, size_t n_floats, int precision, void** out) {
  TFPNP_CHECK(weights_host && out, "null argument");
  TFPNP_CHECK(n_floats == kUnetParamCount, "UNet(2,1) has %zu parameters, got %zu", kUnetParamCount, n_floats);
  int dev = 0;
  TFPNP_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TFPNP_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("tfpnp_b200 needs an sm_100 device (B200); found sm_%d%d", prop.major, prop.minor);
    return TFPNP_ERR_UNSUPPORTED;
  }
  Denoiser* d = nullptr;
  if (precision == TFPNP_PREC_FP32_SIMT) d = make_unet_simt(weights_host);
  else if (precision == TFPNP_PREC_FP16 || precision == TFPNP_PREC_FP16X3) d = make_unet_tc(weights_host, precision);
  else { set_error("unknown precision %d", precision); return TFPNP_ERR_INVALID; }
  if (!d) return TFPNP_ERR_CUDA;
  *out = d;
  return 0;
}

int tfpnp_ircnn_create(const float* weights_host, size_t n_floats, int precision, void** out) {
  TFPNP_CHECK(weights_host && out, "null argument");
  TFPNP_CHECK(n_floats == kIrcnnParamCount, "IRCNN(2,1,64) has %zu parameters, got %zu", kIrcnnParamCount, n_floats);
  int dev = 0;
  TFPNP_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  TFPNP_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("tfpnp_b200 needs an sm_100 device (B200); found sm_%d%d", prop.major, prop.minor);
    return TFPNP_ERR_UNSUPPORTED;
  }
  if (precision != TFPNP_PREC_FP16 && precision != TFPNP_PREC_FP16X3) {
    set_error("IRCNN supports precision fp16 / fp16x3, got %d", precision);
    return TFPNP_ERR_INVALID;
  }
  Denoiser* d = make_ircnn_tc(weights_host, precision);
  if (!d) return TFPNP_ERR_CUDA;
  *out = d;
  return 0;
}

int tfpnp_denoiser_destroy(void* h) {
  delete static_cast<Denoiser*>(h);
  return 0;
}

int tfpnp_denoiser_forward(void* h, const float* x, const float* sigma, int64_t sstride, float* out, int B,
                           int H, int W, void* stream) {
  TFPNP_CHECK(h && x && sigma && out && B > 0, "bad argument");
  Denoiser* d = static_cast<Denoiser*>(h);
  TFPNP_TRY(d->prepare(B, H, W));
  return d->forward(x, sigma, sstride, out, B, H, W, static_cast<cudaStream_t>(stream));
}

int tfpnp_denoiser_vjp(void* h, const float* x, const float* sigma, int64_t sstride, const float* gout, float* gx,
                       float* gsigma, int B, int H, int W, void* stream) {
  TFPNP_CHECK(h && x && sigma && gout && gx && gsigma && B > 0, "bad argument");
  return static_cast<Denoiser*>(h)->vjp(x, sigma, sstride, gout, gx, gsigma, 1, B, H, W, static_cast<cudaStream_t>(stream));
}

int tfpnp_denoiser_layer_profile(void* h, const float* x, const float* sigma, float* out, int B, int H, int W, int reps,
                                 float* ms_out, int cap, int* n_out, char* names, void* stream) {
  TFPNP_CHECK(h && x && sigma && out && ms_out && n_out && names && B > 0 && reps > 0 && cap > 0, "bad argument");
  Denoiser* d = static_cast<Denoiser*>(h);
  TFPNP_TRY(d->prepare(B, H, W));
  return d->layer_profile(x, sigma, out, B, H, W, reps, ms_out, cap, n_out, names, static_cast<cudaStream_t>(stream));
}

int tfpnp_debug_grad_workspace(void* h, float* out_host, size_t n_floats, size_t* have_floats) {
  TFPNP_CHECK(h && have_floats, "bad argument");
  size_t n = 0;
  const float* ws = static_cast<Denoiser*>(h)->grad_workspace(&n);
  *have_floats = n;
  if (!out_host || !ws || n == 0) return 0;
  TFPNP_CUDA_OK(cudaDeviceSynchronize());
  TFPNP_CUDA_OK(cudaMemcpy(out_host, ws, (n < n_floats ? n : n_floats) * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int tfpnp_solver_create(const tfpnp_solver_config* cfg, void* denoiser, void** out) {
  return solver_create(cfg, static_cast<Denoiser*>(denoiser), out);
}
int tfpnp_solver_destroy(void* h) {
  delete static_cast<Solver*>(h);
  return 0;
}
int tfpnp_solver_forward(void* h, const float* state_in, const void* aux0, const void* aux1,
                         int64_t aux1_stride, const float* sigma_d, const float* mu, const float* tau,
                         int64_t row_stride, int64_t col_stride, int B, int iters, float* state_out,
                         void* stream) {
  TFPNP_CHECK(h, "null handle");
  return static_cast<Solver*>(h)->forward(state_in, aux0, aux1, aux1_stride, sigma_d, mu, tau, row_stride,
                                          col_stride, B, iters, state_out, static_cast<cudaStream_t>(stream));
}
int64_t tfpnp_solver_last_launch_count(void* h) { return h ? static_cast<Solver*>(h)->last_launches : -1; }
int tfpnp_solver_set_profiling(void* h, int enable) {
  TFPNP_CHECK(h, "null handle");
  static_cast<Solver*>(h)->profiling = enable;
  return 0;
}
int tfpnp_solver_get_profile(void* h, float* den_ms, float* upd_ms) {
  TFPNP_CHECK(h && den_ms && upd_ms, "null argument");
  return static_cast<Solver*>(h)->get_profile(den_ms, upd_ms);
}

}  // extern "C"
