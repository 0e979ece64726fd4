"""Seeded synthetic weights / measurements for the PnP-ADMM hot path.
TEST + BENCH INPUT GENERATION; not product code.

Everything is generated on the CPU with ``torch.Generator().manual_seed(seed)``
so the CPU oracle and the GPU path see identical bits.  The measurement models
mirror the reference's dataset code (cited per function); the reference's own
images / masks / weights are absent (git-ignored upstream).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch

from . import pnp_oracle as O

SEED = 1234  # the reference's default --seed (tfpnp/utils/options.py:28)


def unet_state_dict(seed: int = 0, init: str = "he", out_scale: float = 0.1, dtype=torch.float32):
    """Seeded UNet(2,1) weights with the reference's state_dict keys/shapes.

    init='default' : U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases,
                     the distribution nn.Conv2d uses (activations decay with depth,
                     the residual is tiny -> a weak numerics test);
    init='he'      : N(0, 2/(1+0.2^2)/fan_in) weights (variance preserving under
                     LeakyReLU(0.2)), small biases, last 1x1 layer scaled by
                     ``out_scale`` so the residual is O(out_scale) like a trained
                     denoiser's -> every layer's numerics matter."""
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for key, shape in O.unet_param_shapes():
        if key.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            if init == "default":
                b = 1.0 / math.sqrt(fan_in)
                w = (torch.rand(shape, generator=g) * 2 - 1) * b
            else:
                w = torch.randn(shape, generator=g) * math.sqrt(2.0 / (1 + 0.04) / fan_in)
                if key.startswith("outc"):
                    w = w * out_scale
            sd[key] = w.to(dtype)
        else:
            fan_in = None
            wkey = key[:-4] + "weight"
            ws = sd[wkey].shape
            fan_in = ws[1] * ws[2] * ws[3]
            if init == "default":
                b = 1.0 / math.sqrt(fan_in)
                sd[key] = ((torch.rand(shape, generator=g) * 2 - 1) * b).to(dtype)
            else:
                sd[key] = (torch.randn(shape, generator=g) * 0.05 *
                           (out_scale if key.startswith("outc") else 1.0)).to(dtype)
    return sd


def ircnn_state_dict(seed: int = 0, out_scale: float = 0.1):
    """Seeded IRCNN(2,1,64) weights (oracle/pnp_oracle.py: ircnn_param_shapes): variance-preserving
    He-uniform for the ReLU layers, the last layer scaled so the predicted residual is a moderate correction."""
    g = torch.Generator().manual_seed(1000 + seed)
    sd = {}
    shapes = O.ircnn_param_shapes()
    for name, shape in shapes:
        if name.endswith("weight"):
            fan_in = shape[1] * 9
            bound = math.sqrt(6.0 / fan_in)
            w = (torch.rand(shape, generator=g) * 2 - 1) * bound
            if name == shapes[-2][0]:
                w = w * out_scale
            sd[name] = w
        else:
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    return sd


def radial_mask(n: int, lines: int) -> torch.Tensor:
    """Boolean union of ``lines`` straight lines through the centre at angles
    k*pi/lines (stand-in for the absent radial_128_{2,4,8}.mat,
    tasks/csmri/main.py:22)."""
    m = torch.zeros(n, n, dtype=torch.bool)
    c = n // 2
    t = torch.arange(-n, n + 1, dtype=torch.float64) / 2.0
    for k in range(lines):
        a = math.pi * k / lines
        ii = torch.round(c + t * math.sin(a)).long()
        jj = torch.round(c + t * math.cos(a)).long()
        ok = (ii >= 0) & (ii < n) & (jj >= 0) & (jj < n)
        m[ii[ok], jj[ok]] = True
    return m


def csmri_batch(B: int, n: int, it: int, seed: int = SEED, sigma_n: float = 15 / 255):
    """tasks/csmri/dataset.py:27-76: y0 = fft2c(gt) + N(0, sigma_n^2), zero off-mask;
    x0 = ifft2c(y0); state = ADMMSolver.reset.  Mask density cycles over the batch."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, n, n, generator=g)
    line_opts = [max(2, n // 3), max(2, n // 6), max(2, n // 12)]
    mask = torch.stack([radial_mask(n, line_opts[b % 3]) for b in range(B)])[:, None]
    y0 = O.fft2c(O.real2complex(gt))
    y0 = y0 + torch.randn(y0.shape, generator=g) * sigma_n
    y0 = y0 * mask[..., None]
    x0 = O.ifft2c(y0)
    sigma_d = torch.rand(B, it, generator=g) * (70 / 255)
    mu = torch.rand(B, it, generator=g)
    return dict(gt=gt, y0=y0, mask=mask, x0=x0, state=O.admm_reset(x0), sigma_d=sigma_d, mu=mu)


def pr_batch(B: int, n: int, it: int, seed: int = SEED, alpha: float = 27.0, n_masks: int = 4):
    """tasks/pr/dataset.py:24-70: 4 unit-modulus CDP masks, y0 = |cdp_forward(gt)| with
    PoissonModel(alpha) noise (tfpnp/utils/noise.py:56-76); x0 = ones."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, n, n, generator=g)
    phi = torch.rand(B, n_masks, n, n, generator=g) * (2 * math.pi)
    mask = torch.stack([torch.cos(phi), torch.sin(phi)], -1)
    z = (O.cdp_forward(O.real2complex(gt), mask) ** 2).sum(-1).sqrt()
    noise = alpha / 255 * z.abs() * torch.randn(z.shape, generator=g)
    y0 = torch.sqrt(torch.clamp(z ** 2 + noise, min=0))
    x0 = torch.ones(B, 1, n, n)
    sigma_d = torch.rand(B, it, generator=g) * (70 / 255)
    mu = torch.rand(B, it, generator=g)
    tau = torch.rand(B, it, generator=g) * 2
    return dict(gt=gt, y0=y0, mask=mask, x0=x0, state=O.pr_reset(x0), sigma_d=sigma_d, mu=mu, tau=tau)


def ct_batch(B: int, n: int, views: int, it: int, seed: int = SEED, noise_p: float = 0.05,
             opnorm: float | None = None):
    """tasks/ct/dataset.py:26-105: y0 = A gt + GaussianModelP(0.05) noise
    (tfpnp/utils/noise.py:36-53); x0 = A^T y0 / opnorm^2; view = full(views/120)."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, n, n, generator=g)
    cs, sn, det = O.ct_geometry(n, views)
    if opnorm is None:
        opnorm = O.radon_opnorm(n, cs, sn, det)
    y0 = O.radon_forward(gt, cs, sn, det)
    y0 = y0 + torch.randn(y0.shape, generator=g) * y0.abs().mean() * noise_p
    x0 = O.radon_backward(y0, cs, sn, n) / opnorm ** 2
    view = torch.full((B, 1, n, n), views / 120.0)
    sigma_d = torch.rand(B, it, generator=g) * (70 / 255)
    mu = torch.rand(B, it, generator=g)
    tau = torch.rand(B, it, generator=g) * 2
    return dict(gt=gt, y0=y0, view=view, x0=x0, state=O.admm_reset(x0), sigma_d=sigma_d, mu=mu,
                tau=tau, opnorm=opnorm, views=views)


def spi_batch(B: int, n: int, it: int, seed: int = SEED):
    """tasks/spi/dataset.py:24-68 + transforms.py:395-401: K cycles {4,6,8};
    binary quanta image y = 1[Poisson(K * kron(gt, 1_KxK) / K^2) >= 1]; x0 = avg_pool(y, K);
    K tensor = full(K/10)."""
    g = torch.Generator().manual_seed(seed)
    gt = torch.rand(B, 1, n, n, generator=g)
    x0 = torch.empty(B, 1, n, n)
    Kt = torch.empty(B, 1, n, n)
    for b in range(B):
        K = (4, 6, 8)[b % 3]
        theta = K * gt[b:b + 1].repeat_interleave(K, 2).repeat_interleave(K, 3) / (K ** 2)
        y = (torch.poisson(theta, generator=g) >= 1).float()
        x0[b] = torch.nn.functional.avg_pool2d(y, K)[0]
        Kt[b] = K / 10.0
    sigma_d = (torch.rand(B, it, generator=g) * 55 + 15) / 255
    mu = torch.rand(B, it, generator=g) * 70 + 50
    return dict(gt=gt, x0=x0, K=Kt, state=O.admm_reset(x0), sigma_d=sigma_d, mu=mu)
