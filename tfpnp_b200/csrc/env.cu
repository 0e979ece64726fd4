// PnPEnv.step bookkeeping on the device (tfpnp/env/base.py:157-191, tasks/*/env.py): the caller side of the
// solver hot path.  The reference re-gathers every state tensor by fancy indexing twice per step
// (`_observation`, base.py:176,188), scatters the solver result with two index_put calls (base.py:171-172) and
// builds the policy observation from six `cat`/`permute`/`contiguous` copies (tasks/csmri/env.py:14-23).
// Here each of these is ONE launch:
//   env_gather        all tensors of an observation, rows idx_left[], 16-byte vectorised
//   env_scatter_state state['solver'][idx] = s  and  state['output'][idx] = get_output(s)  fused
//   env_policy_ob     channel packing (complex2real / complex2channel / bool->float / cat) with the gather fused
#include "common.cuh"

namespace tfpnp {
namespace {

constexpr int kMaxGather = 12;
constexpr int kMaxObChan = 32;   // PR with 8 masks packs 5 + 3*8 = 29 channels (tasks/pr/env.py:10)

struct GatherParams {
  const uint8_t* src[kMaxGather];
  uint8_t* dst[kMaxGather];
  int64_t row_bytes[kMaxGather];
  const int64_t* idx;   // nullptr = identity
};

// grid (chunks, n_rows, n_items); a row is copied in 16-byte words when src, dst and row_bytes allow it
__global__ void __launch_bounds__(256)
env_gather_kernel(const __grid_constant__ GatherParams p) {
  const int t = blockIdx.z;
  const int64_t rb = p.row_bytes[t];
  const int64_t r = blockIdx.y;
  const int64_t sr = p.idx ? p.idx[r] : r;
  const uint8_t* s = p.src[t] + sr * rb;
  uint8_t* d = p.dst[t] + r * rb;
  const bool vec = ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | (uintptr_t)rb) & 15) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (vec) {
    const int64_t n = rb >> 4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < rb; i += stride) d[i] = s[i];
  }
}

// solver_state [n,V,HW(,2)] -> state_solver[idx[r]] (whole row) and state_output[idx[r]] = x (real part);
// V = solver.num_var: 3 for ADMM / iADMM / RED-ADMM, 2 for HQS / APG, 1 for PG (tfpnp/pnp/solver/base.py:87-214)
template <bool COMPLEX>
__global__ void __launch_bounds__(256)
env_scatter_state_kernel(const float* __restrict__ st, const int64_t* __restrict__ idx, float* __restrict__ state_solver,
                         float* __restrict__ state_output, int64_t HW, int num_var) {
  constexpr int E = COMPLEX ? 2 : 1;
  const int64_t r = blockIdx.y;
  const int64_t dr = idx ? idx[r] : r;
  const float* s = st + r * num_var * HW * E;
  float* ds = state_solver + dr * num_var * HW * E;
  float* dout = state_output + dr * HW;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) {
    if (COMPLEX) {
      const float2 x = reinterpret_cast<const float2*>(s)[i];
      reinterpret_cast<float2*>(ds)[i] = x;
      for (int v = 1; v < num_var; ++v)
        reinterpret_cast<float2*>(ds)[v * HW + i] = reinterpret_cast<const float2*>(s)[v * HW + i];
      dout[i] = x.x;                       // get_output: x[..., 0] (tasks/csmri/solver.py:9-18)
    } else {
      const float x = s[i];
      ds[i] = x;
      for (int v = 1; v < num_var; ++v) ds[v * HW + i] = s[v * HW + i];
      dout[i] = x;                         // get_output: first 1/num_var of dim 1 (base.py:101-104)
    }
  }
}

struct ObParams {
  const void* src[kMaxObChan];
  int64_t img_stride[kMaxObChan];   // elements between images of the source tensor
  int64_t offset[kMaxObChan];       // element offset of this channel inside an image
  int32_t pix_stride[kMaxObChan];   // 1 (real plane) or 2 (re / im of an interleaved complex plane)
  int32_t dtype[kMaxObChan];        // 0 = f32, 1 = u8 (torch.bool mask -> float)
  const int64_t* idx;
  float* dst;
  int n_ch;
};

// dst[r, c, i] = float(src_c[idx[r]*img_stride + offset + i*pix_stride]);  grid (chunks, n_ch, n_rows)
__global__ void __launch_bounds__(256)
env_policy_ob_kernel(const __grid_constant__ ObParams p, int64_t HW) {
  const int c = blockIdx.y;
  const int64_t r = blockIdx.z;
  const int64_t sr = p.idx ? p.idx[r] : r;
  float* d = p.dst + (r * p.n_ch + c) * HW;
  const int64_t base = sr * p.img_stride[c] + p.offset[c];
  const int ps = p.pix_stride[c];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  if (p.dtype[c] == 0) {
    const float* s = static_cast<const float*>(p.src[c]) + base;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) d[i] = s[i * ps];
  } else {
    const uint8_t* s = static_cast<const uint8_t*>(p.src[c]) + base;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += stride) d[i] = s[i * ps] ? 1.f : 0.f;
  }
}

}  // namespace
}  // namespace tfpnp

using namespace tfpnp;

extern "C" {

int tfpnp_env_gather(const tfpnp_gather_item* items, int n_items, const int64_t* idx, int n_rows, void* stream) {
  TFPNP_CHECK(items && n_items > 0 && n_items <= kMaxGather, "env_gather: 1..%d tensors per call, got %d", kMaxGather, n_items);
  if (n_rows == 0) return 0;
  TFPNP_CHECK(n_rows > 0, "env_gather: negative row count");
  GatherParams p{};
  int64_t max_rb = 0;
  for (int i = 0; i < n_items; ++i) {
    TFPNP_CHECK(items[i].src && items[i].dst && items[i].row_bytes > 0, "env_gather: bad item %d", i);
    p.src[i] = static_cast<const uint8_t*>(items[i].src);
    p.dst[i] = static_cast<uint8_t*>(items[i].dst);
    p.row_bytes[i] = items[i].row_bytes;
    if (items[i].row_bytes > max_rb) max_rb = items[i].row_bytes;
  }
  p.idx = idx;
  int chunks = (int)((max_rb / 16 + 255) / 256);
  chunks = chunks < 1 ? 1 : (chunks > 64 ? 64 : chunks);
  env_gather_kernel<<<dim3(chunks, n_rows, n_items), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int tfpnp_env_scatter_state(const float* solver_state, const int64_t* idx, int n_rows, float* state_solver,
                            float* state_output, int64_t HW, int complex_state, int num_var, void* stream) {
  TFPNP_CHECK(solver_state && state_solver && state_output && HW > 0 && n_rows >= 0, "env_scatter_state: bad argument");
  TFPNP_CHECK(num_var >= 1 && num_var <= 8, "env_scatter_state: num_var must be 1..8, got %d", num_var);
  if (n_rows == 0) return 0;
  int chunks = (int)((HW + 255) / 256);
  chunks = chunks > 64 ? 64 : chunks;
  dim3 grid(chunks, n_rows);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (complex_state) env_scatter_state_kernel<true><<<grid, 256, 0, st>>>(solver_state, idx, state_solver, state_output, HW, num_var);
  else env_scatter_state_kernel<false><<<grid, 256, 0, st>>>(solver_state, idx, state_solver, state_output, HW, num_var);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

int tfpnp_env_policy_ob(const tfpnp_ob_channel* ch, int n_ch, const int64_t* idx, int n_rows, int64_t HW, float* dst,
                        void* stream) {
  TFPNP_CHECK(ch && dst && n_ch > 0 && n_ch <= kMaxObChan && HW > 0 && n_rows >= 0,
              "env_policy_ob: bad argument (1..%d channels)", kMaxObChan);
  if (n_rows == 0) return 0;
  ObParams p{};
  for (int i = 0; i < n_ch; ++i) {
    TFPNP_CHECK(ch[i].src && (ch[i].pix_stride == 1 || ch[i].pix_stride == 2) && (ch[i].dtype == 0 || ch[i].dtype == 1),
                "env_policy_ob: bad channel %d", i);
    p.src[i] = ch[i].src; p.img_stride[i] = ch[i].img_stride; p.offset[i] = ch[i].offset;
    p.pix_stride[i] = ch[i].pix_stride; p.dtype[i] = ch[i].dtype;
  }
  p.idx = idx; p.dst = dst; p.n_ch = n_ch;
  int chunks = (int)((HW + 255) / 256);
  chunks = chunks > 32 ? 32 : chunks;
  env_policy_ob_kernel<<<dim3(chunks, n_ch, n_rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, HW);
  TFPNP_COUNT_LAUNCH();
  TFPNP_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
