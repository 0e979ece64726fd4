// The other CS-MRI plug-and-play solvers of the reference's `_solver_map` (tasks/csmri/solver.py:60-204):
//   HQS      (:60-88)    x = D(Re z);                         z = ifft2(DC(fft2(x)))
//   PG       (:91-118)   z = x - tau ifft2(M (fft2(x) - y0));  x = D(Re z)
//   APG      (:121-161)  z = s - tau ifft2(M (fft2(s) - y0));  x' = x; x = D(Re z); s = x + beta (x - x')
//   RED-ADMM (:164-201)  x = (lam D(Re x) + mu (z - u)) / (mu + lam);  z = ifft2(DC(fft2(x + u)));  u += x - z
// with DC(Z)[mask] = (mu Z + y0) / (1 + mu) and M the sampling mask.  They share every kernel with the ADMM path: the
// denoiser, the warp-register FFT (fft.cuh) and the pre-rolled / sign-folded k-space operands of csmri_prep; only the
// pointwise step in k-space (blend vs. masked residual) and the AXPY around the transform differ.  Implemented as a
// generic "masked-FFT step" G = ifft2(op(fft2(in))) on complex images (rows -> columns + pointwise -> rows, the same three
// launches as csmri.cu) plus one small pointwise kernel per algorithm.  (AMP, :204-245, draws torch.randn_like inside the
// loop and is not reproducible against a fixture: not built.)
#include "tasks.cuh"
#include "fft.cuh"
#include <map>
#include <cstdlib>

namespace tfpnp {
namespace {

constexpr int V_ROWS_PER_CTA = 8;
constexpr int V_COLS_PER_CTA = 8;
enum { MODE_BLEND = 0, MODE_RESIDUAL = 1 };

template <int R>
__global__ void __launch_bounds__(V_ROWS_PER_CTA * 32)
vstep_rows_fwd(const float2* __restrict__ in, float2* __restrict__ T) {
  constexpr int N = 32 * R;
  WarpFFT<R> f;
  f.init();
  const size_t row = (size_t)blockIdx.x * V_ROWS_PER_CTA + (threadIdx.x >> 5);
  const float2* src = in + row * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = src[32 * j + f.lane];
  f.forward(v);
  float2* dst = T + row * N;
#pragma unroll
  for (int j = 0; j < R; ++j) dst[32 * j + f.lane] = v[j];
}

template <int R, int MODE>
__global__ void __launch_bounds__(V_COLS_PER_CTA * 32)
vstep_cols(float2* __restrict__ T, const float2* __restrict__ y0p, const uint8_t* __restrict__ maskp,
           const float* __restrict__ mu) {
  constexpr int N = 32 * R;
  constexpr int PITCH = V_COLS_PER_CTA + 1;
  __shared__ float2 tile[N * PITCH];
  const int b = blockIdx.y, c0 = blockIdx.x * V_COLS_PER_CTA;
  float2* Tb = T + (size_t)b * N * N;
  for (int i = threadIdx.x; i < N * V_COLS_PER_CTA; i += V_COLS_PER_CTA * 32) {
    const int r = i / V_COLS_PER_CTA, cc = i % V_COLS_PER_CTA;
    tile[r * PITCH + cc] = Tb[(size_t)r * N + c0 + cc];
  }
  __syncthreads();
  WarpFFT<R> f;
  f.init();
  const int w = threadIdx.x >> 5;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = tile[(32 * j + f.lane) * PITCH + w];
  f.forward(v);
  const float m = MODE == MODE_BLEND ? mu[b] : 0.f;
  const float inv_n = 1.0f / (float)N;
  const size_t col = ((size_t)b * N + c0 + w) * N;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float2 zf = make_float2(v[j].x * inv_n, v[j].y * inv_n);
    const bool on = maskp[col + 32 * j + f.lane] != 0;
    const float2 y = y0p[col + 32 * j + f.lane];
    if (MODE == MODE_BLEND) {
      if (on) {                                   // z[mask] = ((mu z + y0)/(1+mu))[mask]   (solver.py:77-79, 191-193)
        zf.x = (m * zf.x + y.x) / (1.0f + m);
        zf.y = (m * zf.y + y.y) / (1.0f + m);
      }
    } else {                                      // temp = fft2(x) - y0; temp[~mask] = 0      (solver.py:108-109, 144-145)
      zf = on ? make_float2(zf.x - y.x, zf.y - y.y) : make_float2(0.f, 0.f);
    }
    v[j] = zf;
  }
  f.inverse(v);
#pragma unroll
  for (int j = 0; j < R; ++j) tile[(32 * j + f.lane) * PITCH + w] = v[j];
  __syncthreads();
  for (int i = threadIdx.x; i < N * V_COLS_PER_CTA; i += V_COLS_PER_CTA * 32) {
    const int r = i / V_COLS_PER_CTA, cc = i % V_COLS_PER_CTA;
    Tb[(size_t)r * N + c0 + cc] = tile[r * PITCH + cc];
  }
}

template <int R>
__global__ void __launch_bounds__(V_ROWS_PER_CTA * 32)
vstep_rows_inv(const float2* __restrict__ T, float2* __restrict__ out) {
  constexpr int N = 32 * R;
  WarpFFT<R> f;
  f.init();
  const size_t row = (size_t)blockIdx.x * V_ROWS_PER_CTA + (threadIdx.x >> 5);
  const float2* src = T + row * N;
  float2 v[R];
#pragma unroll
  for (int j = 0; j < R; ++j) v[j] = src[32 * j + f.lane];
  f.inverse(v);
  const float inv_n = 1.0f / (float)N;
  float2* dst = out + row * N;
#pragma unroll
  for (int j = 0; j < R; ++j) dst[32 * j + f.lane] = make_float2(v[j].x * inv_n, v[j].y * inv_n);
}

template <int R>
int masked_fft_step(const float2* in, float2* G, float2* T, const float2* y0p, const uint8_t* maskp, const float* mu,
                    int mode, int B, cudaStream_t st) {
  constexpr int N = 32 * R;
  const int row_blocks = B * N / V_ROWS_PER_CTA;
  vstep_rows_fwd<R><<<row_blocks, V_ROWS_PER_CTA * 32, 0, st>>>(in, T);
  if (mode == MODE_BLEND)
    vstep_cols<R, MODE_BLEND><<<dim3(N / V_COLS_PER_CTA, B), V_COLS_PER_CTA * 32, 0, st>>>(T, y0p, maskp, mu);
  else
    vstep_cols<R, MODE_RESIDUAL><<<dim3(N / V_COLS_PER_CTA, B), V_COLS_PER_CTA * 32, 0, st>>>(T, y0p, maskp, mu);
  vstep_rows_inv<R><<<row_blocks, V_ROWS_PER_CTA * 32, 0, st>>>(T, G);
  TFPNP_COUNT_LAUNCH(); TFPNP_COUNT_LAUNCH(); TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- pointwise pieces (one thread per pixel; per-image parameters p[b]) -----------------------------------------
// HQS: z = G; d = Re z
__global__ void hqs_finish(const float2* __restrict__ G, float2* __restrict__ z, float* __restrict__ d, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 g = G[i];
  z[i] = g; d[i] = g.x;
}
// x (real, denoiser output) -> complex (x, 0): real2complex (transforms.py:12-13)
__global__ void real_to_complex(const float* __restrict__ x, float2* __restrict__ xc, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xc[i] = make_float2(x[i], 0.f);
}
// PG / APG: z = in - tau G; d = Re z                                              (solver.py:110-111, 146)
__global__ void grad_step(const float2* __restrict__ in, const float2* __restrict__ G, const float* __restrict__ tau,
                          float* __restrict__ d, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = tau[i / HW];
  d[i] = in[i].x - t * G[i].x;          // only Re z feeds the denoiser; z itself is not part of the state
}
// APG: s = x + beta (x - x_prev); x_prev = x   (x real from the denoiser, as complex)   (solver.py:149-158)
__global__ void apg_extrapolate(const float* __restrict__ x, float2* __restrict__ xprev, float2* __restrict__ s,
                                const float* __restrict__ beta, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float bt = beta[i / HW];
  const float2 xc = make_float2(x[i], 0.f), xp = xprev[i];
  s[i] = make_float2(xc.x + bt * (xc.x - xp.x), xc.y + bt * (xc.y - xp.y));
  xprev[i] = xc;
}
// RED-ADMM x step: x = (lam x_half + mu (z - u)) / (mu + lam); in = x + u             (solver.py:184-188)
__global__ void red_xstep(const float* __restrict__ xh, float2* __restrict__ xc, const float2* __restrict__ z,
                          const float2* __restrict__ u, float2* __restrict__ in, const float* __restrict__ mu,
                          const float* __restrict__ lam, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float m = mu[i / HW], l = lam[i / HW];
  const float2 zz = z[i], uu = u[i];
  const float2 xn = make_float2((l * xh[i] + m * (zz.x - uu.x)) / (m + l), (l * 0.f + m * (zz.y - uu.y)) / (m + l));
  xc[i] = xn;
  in[i] = make_float2(xn.x + uu.x, xn.y + uu.y);
}
// RED-ADMM z / u step: z = G; u = u + x - z; d = Re x (the next denoiser input)       (solver.py:190-197, 184)
__global__ void red_finish(const float2* __restrict__ G, const float2* __restrict__ xc, float2* __restrict__ z,
                           float2* __restrict__ u, float* __restrict__ d, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float2 g = G[i], x = xc[i];
  float2 uu = u[i];
  uu.x = uu.x + x.x - g.x;
  uu.y = uu.y + x.y - g.y;
  z[i] = g; u[i] = uu; d[i] = x.x;
}
// slot k of a [B, V, HW] complex state <-> a [B, HW] complex buffer
__global__ void slot_copy(const float2* __restrict__ state, float2* __restrict__ buf, float* __restrict__ re, int V, int k,
                          int HW, size_t n, int to_state) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const size_t b = i / HW, p = i % HW;
  float2* s = const_cast<float2*>(state) + (b * V + k) * HW + p;
  if (to_state) *s = buf[i];
  else { const float2 v = *s; buf[i] = v; if (re) re[i] = v.x; }
}
__global__ void gather_params3(const float* __restrict__ p0, const float* __restrict__ p1, const float* __restrict__ p2,
                               int64_t rs, int64_t cs, float* __restrict__ P, int B, int iters) {
  const int n = B * iters;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = t / B, b = t % B;
    const int64_t src = b * rs + i * cs;
    P[t] = p0[src];
    if (p1) P[n + t] = p1[src];
    if (p2) P[2 * n + t] = p2[src];
  }
}

struct VariantSolver {
  int algo = 0, N = 0;
  Denoiser* den = nullptr;
  int cap_B = 0;
  size_t cap_params = 0;
  DevBuf c0, c1, c2, c3, G, T, xr, d, y0p, maskp, params;   // complex images c0..c3, G, T; real x, d
  int64_t last_launches = 0;
  // the iteration loop of a call is ONE CUDA-graph launch, keyed by (B, iters), like the ADMM solvers (solver.cu): the graph
  // only touches the resident buffers; the caller's tensors are read / written by the copy kernels around it
  bool use_graph = true;
  std::map<std::pair<int, int>, cudaGraphExec_t> graphs;
  std::map<std::pair<int, int>, int64_t> graph_nodes;
  cudaStream_t cap_stream = nullptr;
  int den_generation = -1;

  void drop_graphs() {
    for (auto& g : graphs) cudaGraphExecDestroy(g.second);
    graphs.clear();
  }

  int ensure(int B, int iters) {
    const size_t HW = (size_t)N * N;
    if (B > cap_B) {
      drop_graphs();                       // the graphs bake the workspace addresses
      for (DevBuf* b : {&c0, &c1, &c2, &c3, &G, &T, &y0p}) TFPNP_TRY(b->alloc(B * HW * sizeof(float2)));
      TFPNP_TRY(xr.alloc(B * HW * sizeof(float)));
      TFPNP_TRY(d.alloc(B * HW * sizeof(float)));
      TFPNP_TRY(maskp.alloc(B * HW));
      cap_B = B;
    }
    const size_t need = (size_t)B * (iters > 0 ? iters : 1) * 3 * sizeof(float);
    if (need > cap_params) { drop_graphs(); TFPNP_TRY(params.alloc(need)); cap_params = params.bytes; }
    return 0;
  }

  // the iteration loop on stream `st`: resident buffers only (graph-capturable: no allocation, no synchronisation)
  template <int R>
  int enqueue_loop(int B, int iters, cudaStream_t st) {
    const int HW = N * N;
    const size_t n = (size_t)B * HW;
    const int T256 = 256;
    const unsigned nb = (unsigned)((n + T256 - 1) / T256);
    const float* P = params.as<float>();
    const size_t np = (size_t)B * iters;
    float2 *X = c0.as<float2>(), *Z = c1.as<float2>(), *U = c2.as<float2>(), *IN = c3.as<float2>(), *g = G.as<float2>();
    float *x = xr.as<float>(), *dd = d.as<float>();
    auto step = [&](const float2* in, int mode, const float* mu) {
      return masked_fft_step<R>(in, g, T.as<float2>(), y0p.as<float2>(), maskp.as<uint8_t>(), mu, mode, B, st);
    };
    auto denoise = [&](const float* sig) { return den->forward(dd, sig, 1, x, B, N, N, st); };
#define VLAUNCH(kernel, ...) do { kernel<<<nb, T256, 0, st>>>(__VA_ARGS__); TFPNP_COUNT_LAUNCH(); } while (0)
    for (int i = 0; i < iters; ++i) {
      switch (algo) {
        case TFPNP_ALGO_HQS:              // state (x, z); params (sigma_d, mu)
          TFPNP_TRY(denoise(P + (size_t)i * B));
          VLAUNCH(real_to_complex, x, X, n);
          TFPNP_TRY(step(X, MODE_BLEND, P + np + (size_t)i * B));
          VLAUNCH(hqs_finish, g, Z, dd, n);
          break;
        case TFPNP_ALGO_PG:               // state x; params (sigma_d, tau)
          TFPNP_TRY(step(X, MODE_RESIDUAL, nullptr));
          VLAUNCH(grad_step, X, g, P + np + (size_t)i * B, dd, HW, n);
          TFPNP_TRY(denoise(P + (size_t)i * B));
          VLAUNCH(real_to_complex, x, X, n);
          break;
        case TFPNP_ALGO_APG:              // state (x, s); params (sigma_d, tau, beta); X holds x_prev, U holds s
          TFPNP_TRY(step(U, MODE_RESIDUAL, nullptr));
          VLAUNCH(grad_step, U, g, P + np + (size_t)i * B, dd, HW, n);
          TFPNP_TRY(denoise(P + (size_t)i * B));
          VLAUNCH(apg_extrapolate, x, X, U, P + 2 * np + (size_t)i * B, HW, n);
          break;
        default: {                        // RED-ADMM: state (x, z, u); params (sigma_d, mu, lamda)
          const float* mu = P + np + (size_t)i * B;
          TFPNP_TRY(denoise(P + (size_t)i * B));
          VLAUNCH(red_xstep, x, X, Z, U, IN, mu, P + 2 * np + (size_t)i * B, HW, n);
          TFPNP_TRY(step(IN, MODE_BLEND, mu));
          VLAUNCH(red_finish, g, X, Z, U, dd, n);
          break;
        }
      }
    }
#undef VLAUNCH
    return 0;
  }

  template <int R>
  int run(const float* state_in, const float* y0, const uint8_t* mask, const float* p0, const float* p1, const float* p2,
          int64_t rs, int64_t cs, int B, int iters, float* state_out, cudaStream_t st) {
    const int HW = N * N;
    const size_t n = (size_t)B * HW;
    const int T256 = 256;
    const unsigned nb = (unsigned)((n + T256 - 1) / T256);
    TFPNP_CHECK(algo >= TFPNP_ALGO_HQS && algo <= TFPNP_ALGO_REDADMM, "unknown CS-MRI solver variant %d", algo);
    const int V = algo == TFPNP_ALGO_PG ? 1 : (algo == TFPNP_ALGO_REDADMM ? 3 : 2);
    const float2* sin = reinterpret_cast<const float2*>(state_in);
    float2* sout = reinterpret_cast<float2*>(state_out);
    TFPNP_CUDA_OK(cudaMemcpyAsync(state_out, state_in, n * V * sizeof(float2), cudaMemcpyDeviceToDevice, st));
    if (iters == 0) return 0;                                    // the reference returns the state unchanged
    gather_params3<<<cdiv(B * iters, T256), T256, 0, st>>>(p0, p1, p2, rs, cs, params.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    TFPNP_TRY(csmri_prep(y0, mask, y0p.as<float2>(), maskp.as<uint8_t>(), B, N, st));
    TFPNP_TRY(den->prepare(B, N, N));
    if (den->generation != den_generation) { drop_graphs(); den_generation = den->generation; }   // a denoiser workspace moved
    float2 *X = c0.as<float2>(), *Z = c1.as<float2>(), *U = c2.as<float2>();
    float* dd = d.as<float>();
#define VLAUNCH(kernel, ...) do { kernel<<<nb, T256, 0, st>>>(__VA_ARGS__); TFPNP_COUNT_LAUNCH(); } while (0)
    // 1. the caller's state -> resident buffers
    switch (algo) {
      case TFPNP_ALGO_HQS: VLAUNCH(slot_copy, sin, Z, dd, V, 1, HW, n, 0); break;                   // d = Re z
      case TFPNP_ALGO_PG: VLAUNCH(slot_copy, sin, X, nullptr, V, 0, HW, n, 0); break;
      case TFPNP_ALGO_APG:
        VLAUNCH(slot_copy, sin, X, nullptr, V, 0, HW, n, 0);
        VLAUNCH(slot_copy, sin, U, nullptr, V, 1, HW, n, 0);
        break;
      default:
        VLAUNCH(slot_copy, sin, X, dd, V, 0, HW, n, 0);                                              // d = Re x
        VLAUNCH(slot_copy, sin, Z, nullptr, V, 1, HW, n, 0);
        VLAUNCH(slot_copy, sin, U, nullptr, V, 2, HW, n, 0);
        break;
    }
    // 2. the iteration loop: one graph launch
    if (use_graph) {
      auto key = std::make_pair(B, iters);
      auto it = graphs.find(key);
      if (it == graphs.end()) {
        if (!cap_stream) TFPNP_CUDA_OK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        const int64_t before = g_launch_count;
        TFPNP_CUDA_OK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_loop<R>(B, iters, cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
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
      TFPNP_TRY(enqueue_loop<R>(B, iters, st));
    }
    // 3. resident buffers -> the caller's output state
    switch (algo) {
      case TFPNP_ALGO_HQS:
        VLAUNCH(slot_copy, sout, X, nullptr, V, 0, HW, n, 1);
        VLAUNCH(slot_copy, sout, Z, nullptr, V, 1, HW, n, 1);
        break;
      case TFPNP_ALGO_PG: VLAUNCH(slot_copy, sout, X, nullptr, V, 0, HW, n, 1); break;
      case TFPNP_ALGO_APG:
        VLAUNCH(slot_copy, sout, X, nullptr, V, 0, HW, n, 1);
        VLAUNCH(slot_copy, sout, U, nullptr, V, 1, HW, n, 1);
        break;
      default:
        VLAUNCH(slot_copy, sout, X, nullptr, V, 0, HW, n, 1);
        VLAUNCH(slot_copy, sout, Z, nullptr, V, 1, HW, n, 1);
        VLAUNCH(slot_copy, sout, U, nullptr, V, 2, HW, n, 1);
        break;
    }
#undef VLAUNCH
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  }

  ~VariantSolver() {
    drop_graphs();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    for (DevBuf* b : {&c0, &c1, &c2, &c3, &G, &T, &xr, &d, &y0p, &maskp, &params}) b->release();
  }
};


// ---- reverse mode of the ADMM solver (tfpnp/env/base.py:193-206 under autograd) ------------------------
// The adjoint recursion, its per-element bodies and the iteration sequence live in grad_elem.cuh (host+device, run on the
// CPU by tests/test_grad.py against autograd through the unmodified reference); here: the kernel wrappers and launches.
__global__ void grad_pre(const float2* __restrict__ GZ, const float2* __restrict__ GU, const float2* __restrict__ st_i,
                         const float2* __restrict__ st_n, float2* __restrict__ A, float2* __restrict__ IN, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::admm_pre_elem(i, GZ, GU, st_i, st_n, A, IN, HW);
}
// g_mu[b] = <A, R>_b / (1 + mu[b])^2 ; one CTA per image
__global__ void __launch_bounds__(256)
grad_mu_reduce(const float2* __restrict__ A, const float2* __restrict__ R, const float* __restrict__ mu,
               float* __restrict__ gmu, int64_t stride, int HW) {
  __shared__ float red[256];
  const int b = blockIdx.x;
  const float2* a = A + (size_t)b * HW;
  const float2* r = R + (size_t)b * HW;
  float s = 0.f;
  for (int p = threadIdx.x; p < HW; p += 256) s += a[p].x * r[p].x + a[p].y * r[p].y;
  red[threadIdx.x] = s;
  __syncthreads();
  for (int k = 128; k > 0; k >>= 1) {
    if (threadIdx.x < k) red[threadIdx.x] += red[threadIdx.x + k];
    __syncthreads();
  }
  if (threadIdx.x == 0) { const float m = 1.f + mu[b]; gmu[b * stride] = red[0] / (m * m); }
}
__global__ void grad_mid(const float2* __restrict__ GX, float2* __restrict__ GU, const float2* __restrict__ Q,
                         const float2* __restrict__ st_i, float* __restrict__ gxt, float* __restrict__ v, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::admm_mid_elem(i, GX, GU, Q, st_i, gxt, v, HW);
}
__global__ void grad_post(const float* __restrict__ gv, float2* __restrict__ GX, float2* __restrict__ GZ,
                          float2* __restrict__ GU, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::admm_post_elem(i, gv, GX, GZ, GU);
}

template <int R>
struct AdmmGradOps {
  static constexpr int N = 32 * R;
  static constexpr int HW = N * N;
  static constexpr int T256 = 256;
  Denoiser* den; int B; cudaStream_t st;
  float2 *T, *y0p, *zero; uint8_t* maskp;
  size_t n() const { return (size_t)B * HW; }
  unsigned nb() const { return (unsigned)((n() + T256 - 1) / T256); }
  int slot_get(const float2* state, float2* buf, int k) {
    slot_copy<<<nb(), T256, 0, st>>>(state, buf, nullptr, 3, k, HW, n(), 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int slot_put(float2* state, const float2* buf, int k) {
    slot_copy<<<nb(), T256, 0, st>>>(state, const_cast<float2*>(buf), nullptr, 3, k, HW, n(), 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int pre(const float2* gz, const float2* gu, const float2* st_i, const float2* st_n, float2* A, float2* IN) {
    grad_pre<<<nb(), T256, 0, st>>>(gz, gu, st_i, st_n, A, IN, HW, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int blend(const float2* A, const float* mu_i, float2* Q) {
    return masked_fft_step<R>(A, Q, T, zero, maskp, mu_i, MODE_BLEND, B, st);
  }
  int residual(const float2* IN, float2* Rr) { return masked_fft_step<R>(IN, Rr, T, y0p, maskp, nullptr, MODE_RESIDUAL, B, st); }
  int mu_reduce(const float2* A, const float2* Rr, const float* mu_i, float* gmu, int64_t stride) {
    grad_mu_reduce<<<B, 256, 0, st>>>(A, Rr, mu_i, gmu, stride, HW);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int mid(const float2* gx, float2* gu, const float2* Q, const float2* st_i, float* gxt, float* v) {
    grad_mid<<<nb(), T256, 0, st>>>(gx, gu, Q, st_i, gxt, v, HW, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int den_vjp(const float* v, const float* sg_i, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return den->vjp(v, sg_i, 1, gxt, gv, gsig, stride, B, N, N, st);
  }
  int post(const float* gv, float2* gx, float2* gz, float2* gu) {
    grad_post<<<nb(), T256, 0, st>>>(gv, gx, gz, gu, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
};

template <int R>
int admm_backward(Denoiser* den, const float* states, const float* y0, const uint8_t* mask, const float* sigma_d,
                  const float* mu, int64_t rs, int64_t cs, int B, int iters, const float* grad_out, float* g_sigma,
                  float* g_mu, float* g_state_in, cudaStream_t st) {
  constexpr int N = 32 * R;
  const int HW = N * N;
  const size_t n = (size_t)B * HW;
  PoolBuf gx, gz, gu, A, IN, Q, Rr, T, y0p, zero, gxt, v, gv, maskp, P;
  auto body = [&]() -> int {
    for (PoolBuf* b : {&gx, &gz, &gu, &A, &IN, &Q, &Rr, &T, &y0p, &zero}) TFPNP_TRY(b->alloc(n * sizeof(float2), st));
    for (PoolBuf* b : {&gxt, &v, &gv}) TFPNP_TRY(b->alloc(n * sizeof(float), st));
    TFPNP_TRY(maskp.alloc(n, st));
    TFPNP_TRY(P.alloc((size_t)B * iters * 3 * sizeof(float), st));
    TFPNP_CUDA_OK(cudaMemsetAsync(zero.p, 0, n * sizeof(float2), st));
    gather_params3<<<cdiv(B * iters, 256), 256, 0, st>>>(sigma_d, mu, nullptr, rs, cs, P.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    TFPNP_TRY(csmri_prep(y0, mask, y0p.as<float2>(), maskp.as<uint8_t>(), B, N, st));
    AdmmGradOps<R> ops{den, B, st, T.as<float2>(), y0p.as<float2>(), zero.as<float2>(), maskp.as<uint8_t>()};
    grad_elem::AdmmGradBufs w{gx.as<float2>(), gz.as<float2>(), gu.as<float2>(), A.as<float2>(), IN.as<float2>(),
                              Q.as<float2>(), Rr.as<float2>(), gxt.as<float>(), v.as<float>(), gv.as<float>()};
    TFPNP_TRY(grad_elem::admm_backward_sequence(ops, reinterpret_cast<const float2*>(states), P.as<float>(), B, HW, iters,
                                                reinterpret_cast<const float2*>(grad_out), g_sigma, g_mu,
                                                reinterpret_cast<float2*>(g_state_in), w));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  };
  const int rc = body();
  for (PoolBuf* b : {&gx, &gz, &gu, &A, &IN, &Q, &Rr, &T, &y0p, &zero, &gxt, &v, &gv, &maskp, &P}) b->release();
  return rc;
}


// ---- reverse mode of HQS / PG / APG / RED-ADMM (grad_elem.cuh: variant_backward_sequence) ------------------------------------
__global__ void var_slot_copy(const float2* __restrict__ state, float2* __restrict__ buf, int V, int k, int HW, size_t n, int to_state) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float2* s = const_cast<float2*>(state) + ((i / HW) * V + k) * HW + i % HW;
  if (to_state) *s = buf[i]; else buf[i] = *s;
}
__global__ void var_pre(int algo, const float2* __restrict__ st_i, const float2* __restrict__ st_n, const float2* __restrict__ G1,
                        const float2* __restrict__ G2, float2* __restrict__ A, float2* __restrict__ IN, int V, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::var_pre_elem(algo, i, st_i, st_n, G1, G2, A, IN, V, HW);
}
__global__ void var_mid(int algo, const float2* __restrict__ st_i, const float2* __restrict__ st_n, const float2* __restrict__ A,
                        const float2* __restrict__ IN, const float2* __restrict__ Q, const float2* __restrict__ R,
                        const float* __restrict__ p1, const float* __restrict__ p2, float2* __restrict__ G0, float2* __restrict__ G1,
                        float2* __restrict__ G2, float* __restrict__ gxt, float* __restrict__ v, float* __restrict__ t1,
                        float* __restrict__ t2, int V, int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::var_mid_elem(algo, i, st_i, st_n, A, IN, Q, R, p1, p2, G0, G1, G2, gxt, v, t1, t2, V, HW);
}
__global__ void var_post1(int algo, const float* __restrict__ gv, float2* __restrict__ A, float2* __restrict__ G0,
                          float2* __restrict__ G1, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::var_post1_elem(algo, i, gv, A, G0, G1);
}
__global__ void var_post2(int algo, const float2* __restrict__ A, const float2* __restrict__ Q, const float2* __restrict__ R,
                          const float* __restrict__ p1, float2* __restrict__ G0, float2* __restrict__ G1, float* __restrict__ t1,
                          int HW, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) grad_elem::var_post2_elem(algo, i, A, Q, R, p1, G0, G1, t1, HW);
}
__global__ void __launch_bounds__(256)
var_image_sum(const float* __restrict__ term, float* __restrict__ out, int64_t stride, int HW) {
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

template <int R>
struct VarGradOps {
  static constexpr int N = 32 * R;
  static constexpr int HW = N * N;
  static constexpr int T256 = 256;
  Denoiser* den; int B; cudaStream_t st;
  float2 *T, *y0p, *zero; uint8_t* maskp;
  size_t n() const { return (size_t)B * HW; }
  unsigned nb() const { return (unsigned)((n() + T256 - 1) / T256); }
  int slot_get(const float2* state, float2* buf, int V, int k) {
    var_slot_copy<<<nb(), T256, 0, st>>>(state, buf, V, k, HW, n(), 0);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int slot_put(float2* state, float2* buf, int V, int k) {
    var_slot_copy<<<nb(), T256, 0, st>>>(state, buf, V, k, HW, n(), 1);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int pre(int algo, const float2* st_i, const float2* st_n, const float2* g1, const float2* g2, float2* A, float2* IN, int V) {
    var_pre<<<nb(), T256, 0, st>>>(algo, st_i, st_n, g1, g2, A, IN, V, HW, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int blend0(const float2* A, const float* mu, float2* Q) { return masked_fft_step<R>(A, Q, T, zero, maskp, mu, MODE_BLEND, B, st); }
  int resid(const float2* in, bool with_y0, float2* out) {
    return masked_fft_step<R>(in, out, T, with_y0 ? y0p : zero, maskp, nullptr, MODE_RESIDUAL, B, st);
  }
  int mid(int algo, const float2* st_i, const float2* st_n, const float2* A, const float2* IN, const float2* Q, const float2* Rr,
          const float* p1, const float* p2, float2* g0, float2* g1, float2* g2, float* gxt, float* v, float* t1, float* t2, int V) {
    var_mid<<<nb(), T256, 0, st>>>(algo, st_i, st_n, A, IN, Q, Rr, p1, p2, g0, g1, g2, gxt, v, t1, t2, V, HW, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int den_vjp(const float* v, const float* sg, const float* gxt, float* gv, float* gsig, int64_t stride) {
    return den->vjp(v, sg, 1, gxt, gv, gsig, stride, B, N, N, st);
  }
  int post1(int algo, const float* gv, float2* A, float2* g0, float2* g1) {
    var_post1<<<nb(), T256, 0, st>>>(algo, gv, A, g0, g1, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int post2(int algo, const float2* A, const float2* Q, const float2* Rr, const float* p1, float2* g0, float2* g1, float* t1) {
    var_post2<<<nb(), T256, 0, st>>>(algo, A, Q, Rr, p1, g0, g1, t1, HW, n());
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
  int reduce(const float* term, float* out, int64_t stride) {
    var_image_sum<<<B, 256, 0, st>>>(term, out, stride, HW);
    TFPNP_COUNT_LAUNCH();
    return 0;
  }
};

template <int R>
int variant_backward(int algo, Denoiser* den, const float* states, const float* y0, const uint8_t* mask, const float* p0,
                     const float* p1, const float* p2, int64_t rs, int64_t cs, int B, int iters, const float* grad_out, float* g_p0,
                     float* g_p1, float* g_p2, float* g_state_in, cudaStream_t st) {
  constexpr int N = 32 * R;
  const int HW = N * N;
  const size_t n = (size_t)B * HW;
  PoolBuf cb[10], fb[5], maskp, P;
  auto body = [&]() -> int {
    for (PoolBuf& b : cb) TFPNP_TRY(b.alloc(n * sizeof(float2), st));
    for (PoolBuf& b : fb) TFPNP_TRY(b.alloc(n * sizeof(float), st));
    TFPNP_TRY(maskp.alloc(n, st));
    TFPNP_TRY(P.alloc((size_t)B * iters * 3 * sizeof(float), st));
    TFPNP_CUDA_OK(cudaMemsetAsync(cb[9].p, 0, n * sizeof(float2), st));                 // zero y0
    TFPNP_CUDA_OK(cudaMemsetAsync(P.p, 0, (size_t)B * iters * 3 * sizeof(float), st));
    gather_params3<<<cdiv(B * iters, 256), 256, 0, st>>>(p0, p1, p2, rs, cs, P.as<float>(), B, iters);
    TFPNP_COUNT_LAUNCH();
    TFPNP_TRY(csmri_prep(y0, mask, cb[8].as<float2>(), maskp.as<uint8_t>(), B, N, st));
    VarGradOps<R> ops{den, B, st, cb[7].as<float2>(), cb[8].as<float2>(), cb[9].as<float2>(), maskp.as<uint8_t>()};
    grad_elem::VarGradBufs w{cb[0].as<float2>(), cb[1].as<float2>(), cb[2].as<float2>(), cb[3].as<float2>(), cb[4].as<float2>(),
                             cb[5].as<float2>(), cb[6].as<float2>(), fb[0].as<float>(), fb[1].as<float>(), fb[2].as<float>(),
                             fb[3].as<float>(), fb[4].as<float>()};
    TFPNP_TRY(grad_elem::variant_backward_sequence(ops, algo, reinterpret_cast<const float2*>(states), P.as<float>(), B, HW, iters,
                                                   reinterpret_cast<const float2*>(grad_out), g_p0, g_p1, g_p2,
                                                   reinterpret_cast<float2*>(g_state_in), w));
    TFPNP_CUDA_OK(cudaGetLastError());
    return 0;
  };
  const int rc = body();
  for (PoolBuf& b : cb) b.release();
  for (PoolBuf& b : fb) b.release();
  maskp.release(); P.release();
  return rc;
}

}  // namespace
}  // namespace tfpnp

using namespace tfpnp;

extern "C" {

int tfpnp_csmri_variant_create(int algo, int N, void* denoiser, void** out) {
  TFPNP_CHECK(denoiser && out, "null argument");
  TFPNP_CHECK(algo >= TFPNP_ALGO_HQS && algo <= TFPNP_ALGO_REDADMM, "unknown CS-MRI solver variant %d", algo);
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "FFT tasks support N in {32,64,128,256}, got %d", N);
  VariantSolver* s = new VariantSolver();
  s->algo = algo; s->N = N; s->den = static_cast<Denoiser*>(denoiser);
  const char* e = getenv("TFPNP_VARIANT_GRAPH");       // 0: eager launches (debugging)
  s->use_graph = !e || atoi(e) != 0;
  *out = s;
  return 0;
}

int tfpnp_csmri_variant_destroy(void* h) {
  delete static_cast<VariantSolver*>(h);
  return 0;
}

int tfpnp_csmri_variant_forward(void* h, const float* state_in, const float* y0, const void* mask, const float* p0,
                                const float* p1, const float* p2, int64_t row_stride, int64_t col_stride, int B,
                                int iters, float* state_out, void* stream) {
  TFPNP_CHECK(h && state_in && state_out && state_in != state_out && y0 && mask && B > 0 && iters >= 0, "bad argument");
  VariantSolver* s = static_cast<VariantSolver*>(h);
  TFPNP_CHECK(iters == 0 || (p0 && p1 && (s->algo < TFPNP_ALGO_APG || p2)), "missing hyper-parameter pointer");
  g_launch_count = 0;
  TFPNP_CUDA_OK(fft_tables_init());     // this translation unit's copy of the twiddle table
  TFPNP_TRY(s->ensure(B, iters));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* m8 = static_cast<const uint8_t*>(mask);
  int rc;
  switch (s->N) {
    case 32: rc = s->run<1>(state_in, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, state_out, st); break;
    case 64: rc = s->run<2>(state_in, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, state_out, st); break;
    case 128: rc = s->run<4>(state_in, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, state_out, st); break;
    default: rc = s->run<8>(state_in, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, state_out, st); break;
  }
  s->last_launches = g_launch_count;
  return rc;
}

int tfpnp_csmri_admm_backward(void* denoiser, const float* states, const float* y0, const void* mask,
                              const float* sigma_d, const float* mu, int64_t row_stride, int64_t col_stride, int B, int N,
                              int iters, const float* grad_out, float* grad_sigma_d, float* grad_mu, float* grad_state_in,
                              void* stream) {
  TFPNP_CHECK(denoiser && states && y0 && mask && sigma_d && mu && grad_out && grad_sigma_d && grad_mu && B > 0 && iters > 0,
              "bad argument");
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "FFT tasks support N in {32,64,128,256}, got %d", N);
  g_launch_count = 0;
  TFPNP_CUDA_OK(fft_tables_init());
  Denoiser* den = static_cast<Denoiser*>(denoiser);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* m8 = static_cast<const uint8_t*>(mask);
  switch (N) {
    case 32: return admm_backward<1>(den, states, y0, m8, sigma_d, mu, row_stride, col_stride, B, iters, grad_out, grad_sigma_d, grad_mu, grad_state_in, st);
    case 64: return admm_backward<2>(den, states, y0, m8, sigma_d, mu, row_stride, col_stride, B, iters, grad_out, grad_sigma_d, grad_mu, grad_state_in, st);
    case 128: return admm_backward<4>(den, states, y0, m8, sigma_d, mu, row_stride, col_stride, B, iters, grad_out, grad_sigma_d, grad_mu, grad_state_in, st);
    default: return admm_backward<8>(den, states, y0, m8, sigma_d, mu, row_stride, col_stride, B, iters, grad_out, grad_sigma_d, grad_mu, grad_state_in, st);
  }
}

int tfpnp_csmri_variant_backward(int algo, void* denoiser, const float* states, const float* y0, const void* mask, const float* p0,
                                 const float* p1, const float* p2, int64_t row_stride, int64_t col_stride, int B, int N, int iters,
                                 const float* grad_out, float* grad_p0, float* grad_p1, float* grad_p2, float* grad_state_in,
                                 void* stream) {
  TFPNP_CHECK(denoiser && states && y0 && mask && p0 && p1 && grad_out && grad_p0 && grad_p1 && B > 0 && iters > 0, "bad argument");
  TFPNP_CHECK(algo >= TFPNP_ALGO_HQS && algo <= TFPNP_ALGO_REDADMM, "unknown CS-MRI solver variant %d", algo);
  TFPNP_CHECK(algo < TFPNP_ALGO_APG || (p2 && grad_p2), "missing third hyper-parameter");
  TFPNP_CHECK(N == 32 || N == 64 || N == 128 || N == 256, "FFT tasks support N in {32,64,128,256}, got %d", N);
  g_launch_count = 0;
  TFPNP_CUDA_OK(fft_tables_init());
  Denoiser* den = static_cast<Denoiser*>(denoiser);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* m8 = static_cast<const uint8_t*>(mask);
  switch (N) {
    case 32: return variant_backward<1>(algo, den, states, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, grad_out, grad_p0, grad_p1, grad_p2, grad_state_in, st);
    case 64: return variant_backward<2>(algo, den, states, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, grad_out, grad_p0, grad_p1, grad_p2, grad_state_in, st);
    case 128: return variant_backward<4>(algo, den, states, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, grad_out, grad_p0, grad_p1, grad_p2, grad_state_in, st);
    default: return variant_backward<8>(algo, den, states, y0, m8, p0, p1, p2, row_stride, col_stride, B, iters, grad_out, grad_p0, grad_p1, grad_p2, grad_state_in, st);
  }
}

}  // extern "C"
