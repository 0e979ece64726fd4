/*
 * tfpnp_b200 -- C ABI of the B200-native PnP-ADMM inner solver.
 *
 * The reference (Vandermode/TFPnP) has no FFI layer: its boundary for this path
 * is the Python duck-type `PnPSolver` (tfpnp/pnp/solver/base.py:5-84) as driven
 * by `PnPEnv.step` (tfpnp/env/base.py:157-191).  The Python classes in
 * tfpnp_b200/solver.py keep that interface and marshal raw device pointers into
 * the entry points below through ctypes.  No torch types cross this boundary:
 * plain pointers, sizes and a cudaStream_t (passed as void*).
 *
 * Conventions
 *  - every function returns 0 on success and a negative tfpnp_status on error;
 *    tfpnp_last_error() returns a thread-local message.  Nothing throws.
 *  - all data pointers are DEVICE pointers, borrowed for the duration of the
 *    call (the kernels are enqueued on `stream`; the caller keeps the buffers
 *    alive until the stream has drained, as PyTorch's caching allocator does).
 *  - a handle is bound to the CUDA device current at creation; it is not
 *    thread-safe.  DataParallel-style use = one handle per device.
 *  - fp32 everywhere unless noted; complex tensors are (re,im)-interleaved.
 */
#ifndef TFPNP_B200_H
#define TFPNP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TFPNP_B200_VERSION 100

typedef enum {
  TFPNP_OK = 0,
  TFPNP_ERR_INVALID = -1,      /* bad argument / unsupported shape            */
  TFPNP_ERR_CUDA = -2,         /* a CUDA runtime / driver call failed         */
  TFPNP_ERR_UNSUPPORTED = -3,  /* device is not sm_100 / feature not built    */
  TFPNP_ERR_NOMEM = -4
} tfpnp_status;

/* which inner loop (tasks/<task>/solver.py) */
typedef enum {
  TFPNP_TASK_CSMRI = 0, /* ADMMSolver_CSMRI.forward  tasks/csmri/solver.py:29-57 */
  TFPNP_TASK_PR = 1,    /* IADMMSolver_PR.forward    tasks/pr/solver.py:37-76    */
  TFPNP_TASK_CT = 2,    /* IADMMSolver_CT.forward    tasks/ct/solver.py:17-53    */
  TFPNP_TASK_SPI = 3    /* ADMMSolver_SPI.forward    tasks/spi/solver.py:17-51   */
} tfpnp_task;

/* arithmetic of the denoiser convolutions */
typedef enum {
  TFPNP_PREC_FP16 = 0,     /* tcgen05 kind::f16, fp16 operands, fp32 accumulate (TMEM)  */
  TFPNP_PREC_FP16X3 = 1,   /* tcgen05, split-fp16 (hi,lo) 3-product emulation of fp32    */
  TFPNP_PREC_FP32_SIMT = 2 /* fp32 FFMA on CUDA cores (verification mode, slow)         */
} tfpnp_precision;

int tfpnp_version(void);
const char* tfpnp_last_error(void);
/* The reverse-mode entry points (tfpnp_*_backward) take their scratch from a pool of cached device blocks kept per
 * (device, stream); this frees the pool (the reference has no counterpart: torch's caching allocator,
 * torch.cuda.empty_cache()). */
int tfpnp_release_cached_scratch(void);

/* ---- denoiser: UNetDenoiser2D (tfpnp/pnp/denoiser/base.py:7-32) ------------ */

/* `weights_host`: the 56 tensors of UNet(2,1).state_dict() (tfpnp/pnp/denoiser/
 * models/unet.py:34-47) flattened and concatenated in state_dict order
 * (n_floats must be 11,773,857).  Copied and re-laid-out once; replaces the
 * per-call weight broadcast of DataParallel (tfpnp/policy/sync_batchnorm/
 * replicate.py:50-75). */
int tfpnp_denoiser_create(const float* weights_host, size_t n_floats, int precision,
                          void** out_handle);
/* IRCNN prox_sigma denoiser (BASELINE configs[0]).  Absent from the reference (tfpnp/pnp/__init__.py:5-13 only
 * knows 'unet'): the published 7-layer, 64-channel, dilation 1-2-3-4-3-2-1 network in inference form,
 * wrapped like UNetDenoiser2D: out = clamp(x - net(cat[x, sigma*ones]), 0, 1).  `weights_host`: the 14 tensors
 * model.{0,2,...,12}.{weight,bias} ([64,2,3,3],[64], 5 x ([64,64,3,3],[64]), [1,64,3,3],[1]) flattened in that
 * order (n_floats = 186,433).  precision: TFPNP_PREC_FP16 or TFPNP_PREC_FP16X3.  The handle is a denoiser
 * handle: use it with tfpnp_denoiser_forward / _destroy and tfpnp_solver_create. */
int tfpnp_ircnn_create(const float* weights_host, size_t n_floats, int precision, void** out_handle);
int tfpnp_denoiser_destroy(void* handle);

/* out[b] = clamp(UNet(cat[x[b], sigma[b]]) , 0, 1)   denoiser/base.py:23-32
 * x, out: [B,1,H,W]; sigma: [B] with element stride `sigma_stride` (floats).
 * H, W multiples of 16 (four 2x poolings), H,W >= 16. */
int tfpnp_denoiser_forward(void* handle, const float* x, const float* sigma,
                           int64_t sigma_stride, float* out, int B, int H, int W, void* stream);

/* Reverse mode of the call above (what autograd computes for the reference's denoiser inside
 * PnPEnv.forward, tfpnp/env/base.py:193-206): for a cotangent gout [B,1,H,W] of `out`,
 *   gx [B,1,H,W] = d<out,gout>/dx,   gsigma [B] = d<out,gout>/dsigma   (weights are frozen: denoiser/base.py:19-21).
 * Recomputes the forward pass keeping every activation (fp32).  Only the TFPNP_PREC_FP32_SIMT engine implements it
 * (first correct path; others return TFPNP_ERR_UNSUPPORTED).  STATUS: validated on a B200 against autograd through the unmodified
 * reference (tests/test_grad.py, part of the default -m gpu suite since round 2). */
int tfpnp_denoiser_vjp(void* handle, const float* x, const float* sigma, int64_t sigma_stride, const float* gout,
                       float* gx, float* gsigma, int B, int H, int W, void* stream);
/* Debugging aid for the call above: copies the reverse-mode workspace of the handle's last tfpnp_denoiser_vjp (every layer's
 * activation and the gradient buffers, layout = grad_elem::unet_vjp_workspace_layout in tfpnp_b200/csrc/grad_elem.cuh) to
 * `out_host` (up to n_floats); *have_floats = its size.  tools/grad_layer_check.py compares it region by region with the CPU
 * emulation to find the first layer that deviates. */
int tfpnp_debug_grad_workspace(void* handle, float* out_host, size_t n_floats, size_t* have_floats);

/* Measurement aid (tools/layer_profile.py -> profiles/): runs `reps` denoiser calls eagerly on `stream` with a CUDA event
 * after every kernel launch and returns the mean device time of each launch of one call, in launch order, in ms_out[0..n)
 * (n = *n_out <= cap; `names` receives n NUL-terminated labels of 48 bytes each, e.g. "l05 128->128 @32 pair").  Warm caches,
 * back-to-back launches, no profiler attached; programmatic dependent launch is off inside it, so the figures include each
 * kernel's own prologue.  Tensor-core engines only.  Synchronises the stream. */
int tfpnp_denoiser_layer_profile(void* handle, const float* x, const float* sigma, float* out, int B, int H, int W, int reps,
                                 float* ms_out, int cap, int* n_out, char* names, void* stream);

/* One denoiser layer on its own (kernel-level parity tests): ConvLayer = nn.Conv2d(3x3, pad 1,
 * bias) + LeakyReLU(0.2) (unet.py:8-22) over the channel concatenation of x0 [B,H,W,C0] and the
 * optional x1 [B,H,W,C1] (torch.cat, unet.py:119), NHWC fp16 in/out, tcgen05 implicit GEMM.
 * w_taps: fp16 [9][Cout][C0+C1] (tap = ky*3+kx); bias fp32 [Cout].  C0, C1 multiples of 32;
 * Cout 32, 64 or a multiple of 128. */
int tfpnp_conv3x3_nhwc(const void* x0, int C0, const void* x1, int C1, const void* w_taps,
                       const float* bias, void* out, int B, int H, int W, int Cout, void* stream);

/* ---- solver: PnPSolver.forward (tfpnp/pnp/solver/base.py:21-32) ------------ */

typedef struct {
  int task;          /* tfpnp_task                                             */
  int H, W;          /* image size                                             */
  int n_masks;       /* PR: number of CDP masks (4); else 0                    */
  int views;         /* CT: number of projection angles; else 0                */
  float opnorm;      /* CT: sqrt(lambda_max(A^T A)) (transforms.py:447-472)    */
  const float* ct_cos; /* CT: optional HOST tables cos/sin(angle_v), [views]; NULL = computed */
  const float* ct_sin; /*     internally from linspace(0, 179pi/180, views)         */
  int use_graph;     /* capture the iteration loop in a CUDA graph (0/1)       */
} tfpnp_solver_config;

int tfpnp_solver_create(const tfpnp_solver_config* cfg, void* denoiser, void** out_handle);
int tfpnp_solver_destroy(void* handle);

/* One `solver(inputs, parameters)` call = `iters` inner iterations.
 *  state_in/state_out : variables cat((x,z,u),1): [B,3,H,W,2] (CSMRI, PR) or [B,3,H,W]
 *                       (CT, SPI); contiguous; state_out may not alias state_in.
 *  aux0, aux1         : CSMRI: y0 [B,1,H,W,2] f32, mask [B,1,H,W] u8 (torch.bool)
 *                       PR   : y0 [B,M,H,W] f32,   mask [B,M,H,W,2] f32
 *                       CT   : y0 [B,1,views,det] f32, NULL
 *                       SPI  : x0 [B,1,H,W] f32,   K [B] f32 with stride aux1_stride
 *                              (= K tensor[:,0,0,0], still divided by 10)
 *  sigma_d, mu, tau   : [B,iters] f32, element (b,i) at p[b*row_stride + i*col_stride]
 *                       (tau NULL for CSMRI/SPI)
 */
int tfpnp_solver_forward(void* handle, const float* state_in, const void* aux0,
                         const void* aux1, int64_t aux1_stride, const float* sigma_d,
                         const float* mu, const float* tau, int64_t row_stride,
                         int64_t col_stride, int B, int iters, float* state_out, void* stream);

/* number of kernels the last tfpnp_solver_forward enqueued (graph nodes included) */
int64_t tfpnp_solver_last_launch_count(void* handle);

/* ---- the other CS-MRI solvers of the reference's _solver_map (tasks/csmri/solver.py:60-204) ----
 * Same kernels as the ADMM path (denoiser, warp FFT, pre-rolled k-space operands); only the pointwise step in
 * k-space and the update around the transform differ.  state_in/state_out: cat of V complex variables
 * [B,V,N,N,2]: HQS (x,z) V=2; PG x V=1; APG (x,s) V=2; RED-ADMM (x,z,u) V=3.  y0 [B,1,N,N,2] f32, mask [B,1,N,N] u8.
 * Hyper-parameters p0,p1,p2 ([B,iters], strided like tfpnp_solver_forward):
 *   HQS (sigma_d, mu, NULL); PG (sigma_d, tau, NULL); APG (sigma_d, tau, beta); RED-ADMM (sigma_d, mu, lamda). */
typedef enum { TFPNP_ALGO_HQS = 1, TFPNP_ALGO_PG = 2, TFPNP_ALGO_APG = 3, TFPNP_ALGO_REDADMM = 4 } tfpnp_csmri_algo;
int tfpnp_csmri_variant_create(int algo, int N, void* denoiser, void** out_handle);
int tfpnp_csmri_variant_destroy(void* handle);
int tfpnp_csmri_variant_forward(void* handle, const float* state_in, const float* y0, const void* mask,
                                const float* p0, const float* p1, const float* p2, int64_t row_stride,
                                int64_t col_stride, int B, int iters, float* state_out, void* stream);

/* Reverse mode of ADMMSolver_CSMRI.forward (tasks/csmri/solver.py:29-57) w.r.t. the hyper-parameters, as the
 * reference's actor update needs it (tfpnp/trainer/mddpg/trainer.py:173 through tfpnp/env/base.py:193-206).
 *   states        [iters+1][B,3,N,N,2]: the input state followed by the state after each iteration (recorded by the
 *                 caller with tfpnp_solver_forward(iters = 1) per iteration)
 *   sigma_d, mu   [B,iters] strided like tfpnp_solver_forward;  y0 [B,1,N,N,2] f32;  mask [B,1,N,N] u8
 *   grad_out      [B,3,N,N,2] cotangent of the final state
 *   grad_sigma_d, grad_mu   [B,iters] contiguous (written);  grad_state_in [B,3,N,N,2] or NULL
 * `denoiser` must implement tfpnp_denoiser_vjp.  Synchronises the stream before returning (scratch is per call). */
int tfpnp_csmri_admm_backward(void* denoiser, const float* states, const float* y0, const void* mask,
                              const float* sigma_d, const float* mu, int64_t row_stride, int64_t col_stride, int B,
                              int N, int iters, const float* grad_out, float* grad_sigma_d, float* grad_mu,
                              float* grad_state_in, void* stream);

/* Reverse mode of ADMMSolver_SPI.forward (tasks/spi/solver.py:17-51), same conventions: states [iters+1][B,3,H,W] real,
 * x0 [B,1,H,W], K [B] with element stride K_stride (the reference's K tensor, i.e. K/10).  Only the closed-form branch of
 * spi_inverse (K1 == 0, transforms.py:415) carries a gradient: the reference's bisection iterates are constants under
 * autograd. */
int tfpnp_spi_admm_backward(void* denoiser, const float* states, const float* x0, const float* K, int64_t K_stride,
                            const float* sigma_d, const float* mu, int64_t row_stride, int64_t col_stride, int B, int H,
                            int W, int iters, const float* grad_out, float* grad_sigma_d, float* grad_mu,
                            float* grad_state_in, void* stream);

/* Reverse mode of IADMMSolver_CT.forward (tasks/ct/solver.py:17-53), same conventions: states [iters+1][B,3,N,N] real,
 * y0 [B,1,views,ceil(sqrt(2)N)], the geometry arguments of tfpnp_radon_forward, opnorm as given to the solver;
 * additionally grad_tau [B,iters].  Uses A^T A = (A^T A)^T (the backprojector is the exact transpose of the projector).
 * Like the CT forward path, pinned to this build's own Radon pair only. */
int tfpnp_ct_iadmm_backward(void* denoiser, const float* states, const float* y0, int views, float opnorm,
                            const float* cos_host, const float* sin_host, const float* sigma_d, const float* mu,
                            const float* tau, int64_t row_stride, int64_t col_stride, int B, int N, int iters,
                            const float* grad_out, float* grad_sigma_d, float* grad_mu, float* grad_tau,
                            float* grad_state_in, void* stream);

/* Reverse mode of IADMMSolver_PR.forward (tasks/pr/solver.py:37-76), same conventions: states [iters+1][B,3,N,N,2],
 * y0 [B,M,N,N] f32, mask [B,M,N,N,2] f32 (unit-modulus CDP masks), M = n_masks.  The magnitude projection
 * h(w) = (1 - y0/|w|) w has a symmetric real Jacobian, so the adjoint of the gradient operator is the operator itself with h
 * replaced by that Jacobian (pr.cu).  FFTs through tfpnp_fft2. */
int tfpnp_pr_iadmm_backward(void* denoiser, const float* states, const float* y0, const float* mask, int n_masks,
                            const float* sigma_d, const float* mu, const float* tau, int64_t row_stride,
                            int64_t col_stride, int B, int N, int iters, const float* grad_out, float* grad_sigma_d,
                            float* grad_mu, float* grad_tau, float* grad_state_in, void* stream);

/* Reverse mode of tfpnp_csmri_variant_forward (HQS / PG / APG / RED-ADMM, tasks/csmri/solver.py:60-201), same conventions:
 * states [iters+1][B,V,N,N,2] recorded with iters = 1 calls; p0,p1,p2 as in the forward; grad_p0..2 [B,iters] contiguous
 * (grad_p2 NULL for the two-parameter solvers); grad_state_in [B,V,N,N,2] or NULL. */
int tfpnp_csmri_variant_backward(int algo, void* denoiser, const float* states, const float* y0, const void* mask,
                                 const float* p0, const float* p1, const float* p2, int64_t row_stride,
                                 int64_t col_stride, int B, int N, int iters, const float* grad_out, float* grad_p0,
                                 float* grad_p1, float* grad_p2, float* grad_state_in, void* stream);

/* ---- CT operators (own discretisation of the reference geometry,
 *      tfpnp/utils/transforms.py:465-491) ------------------------------------- */
/* img [B,1,N,N] <-> sino [B,1,views,ceil(sqrt(2)N)]; cos/sin: optional HOST tables as above */
int tfpnp_radon_forward(const float* img, float* sino, int B, int N, int views,
                        const float* cos_host, const float* sin_host, void* stream);
int tfpnp_radon_backward(const float* sino, float* img, int B, int N, int views,
                         const float* cos_host, const float* sin_host, void* stream);

/* ---- stand-alone transforms (tfpnp/utils/transforms.py:68-103, 282-320) ------
 * out = FFT2 (inverse = 0) or IFFT2 (inverse = 1), ortho-normalised, of n_imgs complex images
 * [n_imgs, N, N, 2] (fp32 interleaved), over the two image dims.  centered = 1 is the reference's
 * fft2 / ifft2 (fftshift(FFT(ifftshift(x)))); centered = 0 is the plain pair used by
 * cdp_forward / cdp_backward.  N in {32, 64, 128, 256}.  workspace: n_imgs*N*N*2 floats; in, out
 * and workspace must not alias.  (The solvers' FFTs are fused with their data-fidelity steps; this
 * operator serves the measurement synthesis and the per-transform parity tests.) */
int tfpnp_fft2(const float* in, float* out, float* workspace, int n_imgs, int N, int inverse,
               int centered, void* stream);

/* ---- reward metric: torch_psnr (tfpnp/env/base.py:237-242) ------------------
 * psnr[b] = 10 log10(1 / mean((clamp(out[b],0,1) - gt[b])^2)); out, gt: [B,HW] */
int tfpnp_psnr(const float* out, const float* gt, float* psnr, int B, int64_t HW, void* stream);
/* Reverse mode of tfpnp_psnr (the reward is part of the actor loss, tfpnp/trainer/mddpg/trainer.py:189):
 * grad_out[b,p] = grad_psnr[b] * d psnr[b] / d out[b,p]; `psnr` is the forward result. */
int tfpnp_psnr_backward(const float* out, const float* gt, const float* psnr, const float* grad_psnr, float* grad_out, int B,
                        int64_t HW, void* stream);

/* ---- environment bookkeeping: PnPEnv.step (tfpnp/env/base.py:157-191) ----------
 * The caller side of the solver on every episode step.  `idx` is a DEVICE vector of int64 row
 * indices (the env's idx_left, base.py:154,181); NULL means the identity.  `items` / `ch` are HOST
 * arrays (copied into the kernel parameters). */

/* dst_t[r] = src_t[idx[r]] for every tensor t of an observation in ONE launch -- replaces the
 * per-tensor fancy indexing of `_observation` (tasks/csmri/env.py:49-57).  Rows are raw bytes. */
typedef struct { const void* src; void* dst; int64_t row_bytes; } tfpnp_gather_item;
int tfpnp_env_gather(const tfpnp_gather_item* items, int n_items, const int64_t* idx, int n_rows,
                     void* stream);

/* state['solver'][idx] = solver_state; state['output'][idx] = solver.get_output(solver_state)
 * (base.py:171-172 with base.py:101-104 / tasks/csmri/solver.py:9-18) fused.
 * solver_state [n_rows,V,HW(,2)]; state_solver [B,V,HW(,2)]; state_output [B,1,HW]; V = num_var is the
 * solver's variable count (3 ADMM / iADMM / RED-ADMM, 2 HQS / APG, 1 PG: tfpnp/pnp/solver/base.py:87-214). */
int tfpnp_env_scatter_state(const float* solver_state, const int64_t* idx, int n_rows,
                            float* state_solver, float* state_output, int64_t HW, int complex_state,
                            int num_var, void* stream);

/* get_policy_ob (tasks/{csmri,pr,ct,spi}/env.py get_policy_ob): dst [n_rows, n_ch, HW] fp32 with
 * dst[r,c,i] = float(src_c[idx[r]*img_stride_c + offset_c + i*pix_stride_c]).  complex2real /
 * complex2channel (transforms.py:16-26) are pix_stride 2 with offset 0 / 1; dtype 1 reads a
 * torch.bool / uint8 plane (mask.float()). */
typedef struct { const void* src; int64_t img_stride; int64_t offset; int32_t pix_stride; int32_t dtype; } tfpnp_ob_channel;
int tfpnp_env_policy_ob(const tfpnp_ob_channel* ch, int n_ch, const int64_t* idx, int n_rows,
                        int64_t HW, float* dst, void* stream);

/* ---- multi-GPU: the one exchange of the data path ------------------------------
 * env_batch shards over ranks with no data-path collective; after the last iteration the per-image PSNR vectors
 * (tfpnp/env/base.py:237-242) are all-gathered.  Replaces DataParallelWithCallback's gather
 * (tfpnp/policy/sync_batchnorm/replicate.py:50-75).  NCCL is resolved at run time (dlopen "libnccl.so.2"); the Python host
 * can use torch.distributed instead (tfpnp_b200/dist.py does by default).
 *   tfpnp_comm_unique_id      rank 0 creates the 128-byte ncclUniqueId; the host broadcasts it by any means
 *   tfpnp_comm_init           every rank, with its device current: communicator of `world` ranks
 *   tfpnp_comm_allgather_psnr out[world * n_local] = concat over ranks of local[n_local] (equal shards), on `stream` */
int tfpnp_comm_unique_id(void* id_out, size_t id_bytes);
int tfpnp_comm_init(const void* id, size_t id_bytes, int rank, int world, void** comm_out);
int tfpnp_comm_destroy(void* comm);
int tfpnp_comm_allgather_psnr(void* comm, const float* local, int n_local, float* out, void* stream);

/* ---- introspection used by tests / bench ---------------------------------- */
/* measured device time (ms) of the denoiser part and the data-fidelity part of the last forward.
 * tfpnp_solver_set_profiling(handle, mode): 0 off; 1 = eager launches with CUDA events between the segments (no graph:
 * slower than the shipped path); 2 = the events are recorded by event nodes INSIDE the captured graph, i.e. the timed
 * launch is the same single graph launch as the product path plus 2 event nodes per iteration (what bench.py reports). */
int tfpnp_solver_set_profiling(void* handle, int enable);
int tfpnp_solver_get_profile(void* handle, float* denoiser_ms, float* update_ms);

#ifdef __cplusplus
}
#endif
#endif /* TFPNP_B200_H */
