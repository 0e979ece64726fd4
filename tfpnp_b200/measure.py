"""Measurement synthesis on the GPU: the forward models of the task datasets, batched, on device.

The reference builds every training / evaluation sample on the CPU inside the DataLoader
(tasks/{csmri,pr,ct,spi}/dataset.py) -- one image at a time, with cuFFT-free torch calls -- and ships the batch to
the GPU in ``PnPEnv.reset``.  Here the same dictionaries (the keys ``PnPEnv.reset`` and the task environments read) are
produced for a whole batch from ground-truth images that already live on the GPU, with the transforms of this package
(``fft2 / ifft2 / cdp_forward``: libtfpnp_b200's FFT operator; ``radon_forward / radon_backward``).  Noise comes from
``torch`` generators on the device (the reference uses the global CPU RNG), so samples are reproducible per seed but
not bit-equal to a CPU-loader run; with the noise switched off the outputs match the reference formulas
(tests/test_measure.py).
"""
from __future__ import annotations

import torch

from . import ops


def _randn_like(x, generator):
    return torch.randn(x.shape, device=x.device, dtype=x.dtype, generator=generator)


def radial_mask(n: int, lines: int, device=None) -> torch.Tensor:
    """A radial k-space sampling mask [n, n] (bool): the union of `lines` straight lines through the centre at angles
    k*pi/lines -- the role of the reference's (undistributed) radial_128_{2,4,8}.mat masks (tasks/csmri/main.py:22)."""
    c = (n - 1) / 2.0
    t = torch.linspace(-c * 1.5, c * 1.5, 4 * n, device=device)
    m = torch.zeros(n, n, dtype=torch.bool, device=device)
    for k in range(lines):
        a = torch.tensor(k * 3.141592653589793 / lines, device=device)
        i = torch.round(c + t * torch.sin(a)).long()
        j = torch.round(c + t * torch.cos(a)).long()
        ok = (i >= 0) & (i < n) & (j >= 0) & (j < n)
        m[i[ok], j[ok]] = True
    return m


def csmri_measure(gt: torch.Tensor, mask: torch.Tensor, sigma_n: float = 0.0, generator=None) -> dict:
    """tasks/csmri/dataset.py:52-66.  gt [B,1,N,N] in [0,1]; mask [B,1,N,N] bool (or broadcastable [1,1,N,N]);
    sigma_n = noise level / 255 already applied (GaussianModelD, tfpnp/utils/noise.py:19-33)."""
    B = gt.shape[0]
    mask = mask.to(gt.device).bool().expand(B, 1, gt.shape[2], gt.shape[3]).contiguous()
    target = gt.float()
    y0 = ops.fft2(torch.stack([target, torch.zeros_like(target)], dim=-1))          # dataset.py:54
    if sigma_n > 0:
        y0 = y0 + _randn_like(y0, generator) * sigma_n                              # noise.py:31
    y0 = y0 * mask[..., None]                                                        # y0[:, ~mask, :] = 0  (dataset.py:59)
    ATy0 = ops.ifft2(y0)
    return {'y0': y0, 'x0': ATy0.clone(), 'ATy0': ATy0, 'gt': target, 'mask': mask,
            'sigma_n': torch.ones_like(y0) * sigma_n, 'output': ATy0[..., 0].clone(), 'input': ATy0.clone()}


def pr_measure(gt: torch.Tensor, mask: torch.Tensor, alpha: float = 0.0, generator=None) -> dict:
    """tasks/pr/dataset.py:49-66.  gt [B,1,N,N]; mask [B,M,N,N,2] unit-modulus CDP masks; alpha: PoissonModel level
    (tfpnp/utils/noise.py:56-76; 0 = noiseless)."""
    target = gt.float()
    Az = ops.cdp_forward(torch.stack([target, torch.zeros_like(target)], dim=-1), mask.float())
    z = (Az ** 2).sum(dim=-1).sqrt()                                                 # complex_abs (transforms.py:106-118)
    sigma = torch.zeros((), device=gt.device)
    y0 = z
    if alpha > 0:
        intensity_noise = alpha / 255 * z.abs() * _randn_like(z, generator)
        y0 = torch.sqrt(torch.clamp(z ** 2 + intensity_noise, min=0))
        sigma = (y0 - z.abs()).std()
    x0 = torch.ones_like(target)
    return {'y0': y0, 'x0': x0, 'output': x0.clone(), 'gt': target, 'mask': mask.float(),
            'sigma_n': torch.ones_like(target) * sigma}


def spi_measure(gt: torch.Tensor, K: int, generator=None) -> dict:
    """tasks/spi/dataset.py:47-63 + transforms.spi_forward (transforms.py:395-401): K x K binary quanta per pixel,
    alpha = K^2, threshold q = 1; x0 = average of the K x K block."""
    target = gt.float()
    theta = (K ** 2) * target.repeat_interleave(K, 2).repeat_interleave(K, 3) / (K ** 2)     # alpha * kron(x, 1_KxK) / K^2
    y = (torch.poisson(theta, generator=generator) >= 1).float()
    x0 = torch.nn.functional.avg_pool2d(y, K)
    return {'x0': x0, 'output': x0.clone(), 'gt': target, 'K': torch.ones_like(target) * K / 10}


def ct_measure(gt: torch.Tensor, views: int, opnorm: float, noise_p: float = 0.0, generator=None) -> dict:
    """tasks/ct/dataset.py:54-104 on this build's Radon pair: y0 = A gt (+ GaussianModelP noise, noise.py:36-53);
    the initial image is A^T y0 / opnorm^2 (the reference's ramp-filtered FBP lives in the absent torch_radon)."""
    target = gt.float()
    y0 = ops.radon_forward(target, views)
    if noise_p > 0:
        y0 = y0 + _randn_like(y0, generator) * y0.abs().mean() * noise_p
    ATy0 = ops.radon_backward(y0, target.shape[-1], views) / opnorm ** 2
    return {'y0': y0, 'x0': ATy0.clone(), 'ATy0': ATy0, 'gt': target, 'view': torch.ones_like(target) * views / 120,
            'output': ATy0.clone(), 'sigma_n': torch.ones_like(target) * noise_p}
