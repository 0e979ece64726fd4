// Per-task data-fidelity kernels (launch wrappers).  Internal layout, per image:
//   x  [H,W]   fp32 real   (denoiser output; the state's x has zero imaginary part)
//   z,u [H,W,2] fp32 complex for CSMRI/PR, [H,W] real for CT/SPI
//   d  [H,W]   fp32 real   next denoiser input Re(z - u)
// Parameters are stored transposed, [iters][B], so one iteration's slice is contiguous.
#pragma once
#include "common.cuh"

namespace tfpnp {

// ---- CS-MRI (tasks/csmri/solver.py:43-55) ------------------------------------
// y0p/maskp: y0 and mask pre-rolled (fftshift), sign-folded and permuted into the FFT's
// native frequency order, stored transposed [B][col][row].
int csmri_prep(const float* y0, const uint8_t* mask, float2* y0p, uint8_t* maskp, int B, int N,
               cudaStream_t st);
// z = ifft2c(DC(fft2c(x + u))); u += x - z; d = Re(z - u)     (3 launches)
int csmri_update(const float* x, float2* z, float2* u, float* d, float2* T, const float2* y0p,
                 const uint8_t* maskp, const float* mu, int B, int N, cudaStream_t st);

// ---- SPI (tasks/spi/solver.py:35-47, transforms.py:404-439) ------------------
// z = spi_inverse(x + u, K1, K, mu); u += x - z; d = z - u      (1 launch)
int spi_update(const float* x, float* z, float* u, float* d, const float* x0, const float* K10,
               const float* mu, int B, int HW, cudaStream_t st);

// ---- PR (tasks/pr/solver.py:50-72, transforms.py:282-320) --------------------
// y0p: |y0| permuted to FFT order, transposed [B][M][col][row]
int pr_prep(const float* y0, float* y0p, int B, int M, int N, cudaStream_t st);
int pr_update(const float* x, float2* z, float2* u, float* d, float2* T, const float* y0p,
              const float2* mask, const float* mu, const float* tau, int B, int M, int N,
              cudaStream_t st);

// ---- CT (tasks/ct/solver.py:32-49) -------------------------------------------
struct CtGeom {
  int N = 0, views = 0, det = 0;
  DevBuf cs, sn;  // fp32 cos/sin tables [views]
  bool bins_always_inside = false;                           // set_tables: unit-norm rows and det >= sqrt(2) N (ct.cu)
  mutable DevBuf tbuf;                                       // zero-padded image + transposed image scratch, 2 x [B,N,N+4]
  size_t scratch_bytes(int B) const;
  int reserve(int B) const;                                  // size tbuf (never during graph capture)
  int init(int N, int views);                                // tables as torch.linspace would give
  int set_tables(const float* cos_host, const float* sin_host);  // caller-supplied tables [views]
};
int radon_forward(const CtGeom& g, const float* img, const float* y0 /*nullable: subtract*/,
                  float* sino, int B, cudaStream_t st);
int radon_backward(const CtGeom& g, const float* sino, float* img, int B, cudaStream_t st);
// z -= tau (A^T(Az - y0)/opnorm^2 + mu (z - (x+u))); u += x - z; d = z - u   (2 launches)
int ct_update(const CtGeom& g, const float* x, float* z, float* u, float* d, float* resid,
              const float* y0, float inv_opnorm2, const float* mu, const float* tau, int B,
              cudaStream_t st);

}  // namespace tfpnp
