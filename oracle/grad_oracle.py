"""Reverse-mode derivatives of the CS-MRI ADMM path.  TEST INFRASTRUCTURE ONLY.

The reference differentiates ``PnPEnv.forward`` (tfpnp/env/base.py:193-206) through
``ADMMSolver_CSMRI.forward`` (tasks/csmri/solver.py:29-57) with PyTorch autograd in the actor update
(tfpnp/trainer/mddpg/trainer.py:158-214): the loss reaches the policy through the hyper-parameters
``sigma_d`` / ``mu`` only.  Two restatements live here:

* ``*_autograd``: autograd through the CPU oracle (``pnp_oracle``), pinned against autograd through the
  UNMODIFIED reference classes by ``oracle/make_golden_grad.py`` (tests/golden/grad_csmri_small.npz);
* ``admm_csmri_vjp_manual``: the hand-derived adjoint recursion, iteration by iteration, in exactly the
  structure the CUDA implementation uses (csmri_variants.cu: ``admm_backward``), so the derivation itself is
  checked on the CPU against autograd.

Adjoint of one iteration  (x', z', u') = step(z, u; sigma, mu), with incoming (gx', gz', gu'):
    gzt = gz' - gu'                               (u' = u + x' - z')
    q   = ifft2c(B_mu fft2c(gzt))                 B_mu = mu/(1+mu) on the sampled set, 1 elsewhere: the k-space blend with
                                                  y0 = 0; F is unitary and B_mu real-diagonal, so the step is self-adjoint
    r   = ifft2c(M (fft2c(x' + u) - y0))          the masked residual of the forward operand
    g_mu    = <gzt, r> / (1 + mu)^2               d/dmu [(mu Z + y0)/(1 + mu)] = (Z - y0)/(1 + mu)^2 on the sampled set
    gxt     = Re(gx' + gu' + q)                   x' is real (real2complex)
    (gv, g_sigma) = J_D(v, sigma)^T gxt           v = Re(z - u): the denoiser's vector-Jacobian product
    gz = (gv, 0);  gu = gu' + q - (gv, 0);  gx = 0
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import pnp_oracle as O

Tensor = torch.Tensor


def denoise_vjp_autograd(sd, x: Tensor, sigma: Tensor, gout: Tensor) -> Tuple[Tensor, Tensor]:
    """(d<out,gout>/dx, d<out,gout>/dsigma) of UNetDenoiser2D.forward (denoiser/base.py:23-32)."""
    x = x.detach().clone().requires_grad_(True)
    sigma = sigma.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = O.denoise(sd, x, sigma)
        gx, gs = torch.autograd.grad(out, (x, sigma), gout)
    return gx, gs


def admm_csmri_vjp_autograd(sd, state, y0, mask, sigma_d, mu, gout):
    """Gradients of <ADMMSolver_CSMRI.forward(...), gout> w.r.t. (sigma_d, mu, state)."""
    state = state.detach().clone().requires_grad_(True)
    sigma_d = sigma_d.detach().clone().requires_grad_(True)
    mu = mu.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = O.admm_csmri(sd, state, y0, mask, sigma_d, mu)
        gs, gm, gst = torch.autograd.grad(out, (sigma_d, mu, state), gout)
    return gs, gm, gst


def admm_csmri_trajectory(sd, state, y0, mask, sigma_d, mu):
    """[state_0, ..., state_it]: what the product records by calling the forward one iteration at a time."""
    states = [state]
    with torch.no_grad():
        for i in range(sigma_d.shape[-1]):
            states.append(O.admm_csmri(sd, states[-1], y0, mask, sigma_d[:, i:i + 1], mu[:, i:i + 1]))
    return states


def admm_csmri_vjp_manual(sd, states, y0, mask, sigma_d, mu, gout, denoise_vjp=denoise_vjp_autograd):
    """The adjoint recursion of the module docstring over a recorded trajectory."""
    B = gout.shape[0]
    it = sigma_d.shape[-1]
    m = mask.bool()[..., None].expand_as(y0)
    gx, gz, gu = (t.clone() for t in O._split3(gout))
    g_sigma = torch.zeros(B, it, dtype=gout.dtype)
    g_mu = torch.zeros(B, it, dtype=gout.dtype)
    zero = torch.zeros_like(y0)
    for i in reversed(range(it)):
        _, z, u = O._split3(states[i])
        xn = O._split3(states[i + 1])[0]
        _mu = mu[:, i].reshape(B, 1, 1, 1, 1)
        gzt = gz - gu
        q = O.ifft2c(O._dc_blend(O.fft2c(gzt), zero, m, _mu))
        r = O.ifft2c(torch.where(m, O.fft2c(xn + u) - y0, zero))
        g_mu[:, i] = (gzt * r).reshape(B, -1).sum(1) / (1 + mu[:, i]) ** 2
        gxt = (gx + gu + q)[..., 0]
        gv, gs = denoise_vjp(sd, O.complex2real(z - u), sigma_d[:, i], gxt)
        g_sigma[:, i] = gs
        gvc = O.real2complex(gv)
        gu = gu + q - gvc
        gz = gvc
        gx = torch.zeros_like(gx)
    return g_sigma, g_mu, torch.cat((gx, gz, gu), dim=1)


# ----------------------------------------------------------------------------
# The denoiser's VJP, layer by layer, in the structure of UNetSimt::vjp (tfpnp_b200/csrc/unet_simt.cu)
# ----------------------------------------------------------------------------

_SPECS = [(2, 32, 0), (32, 32, 0), (32, 32, 0), (32, 64, 1), (64, 64, 1), (64, 64, 1), (64, 128, 2), (128, 128, 2),
          (128, 128, 2), (128, 256, 3), (256, 256, 3), (256, 256, 3), (256, 512, 4), (512, 512, 4), (512, 512, 4),
          (768, 256, 3), (256, 256, 3), (256, 256, 3), (384, 128, 2), (128, 128, 2), (128, 128, 2), (192, 64, 1),
          (64, 64, 1), (64, 64, 1), (96, 32, 0), (32, 32, 0), (32, 32, 0)]
_CH = [32, 64, 128, 256, 512]


def _layer_keys():
    names = ["inc.conv"] + [f"down{i}.mpconv.1" for i in range(1, 5)] + [f"up{i}.conv" for i in range(1, 5)]
    return [f"{n}.conv-{k}.conv2d" for n in names for k in range(3)]


def _up_matrix(h: int) -> Tensor:
    """[2h, h] interpolation matrix of nn.Upsample(x2, bilinear, align_corners=True) with the float expressions of the
    CUDA kernels (upsample2_simt / up_weight)."""
    Ho = 2 * h
    s = torch.tensor(float(h - 1), dtype=torch.float32) / torch.tensor(float(Ho - 1), dtype=torch.float32)
    M = torch.zeros(Ho, h, dtype=torch.float32)
    for Y in range(Ho):
        f = s * torch.tensor(float(Y), dtype=torch.float32)
        y0 = int(f.item())
        y1 = y0 + (1 if y0 < h - 1 else 0)
        l = f - y0
        M[Y, y0] += 1 - l
        M[Y, y1] += l
    return M


def _lrelu_d(a: Tensor) -> Tensor:
    return torch.where(a > 0, torch.ones_like(a), torch.full_like(a, 0.2))


def _pool_bwd(gpool: Tensor, a: Tensor) -> Tensor:
    """Adjoint of MaxPool2d(2): the first maximum in scan order receives the gradient."""
    B, C, H, W = a.shape
    v = a.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)
    arg = torch.zeros(v.shape[:-1], dtype=torch.long)
    best = v[..., 0].clone()
    for k in range(1, 4):
        better = v[..., k] > best
        arg = torch.where(better, torch.full_like(arg, k), arg)
        best = torch.where(better, v[..., k], best)
    g = torch.zeros_like(v)
    g.scatter_(-1, arg[..., None], gpool[..., None])
    return g.reshape(B, C, H // 2, W // 2, 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H, W)


def denoise_vjp_manual(sd, x: Tensor, sigma: Tensor, gout: Tensor) -> Tuple[Tensor, Tensor]:
    import torch.nn.functional as F
    keys = _layer_keys()
    B, _, H, W = x.shape
    with torch.no_grad():
        # forward, every activation kept
        in2 = torch.cat([x, torch.ones_like(x) * sigma.reshape(B, 1, 1, 1)], dim=1)
        a = [None] * 27

        def conv(l, inp):
            return F.leaky_relu(F.conv2d(inp, sd[keys[l] + ".weight"], sd[keys[l] + ".bias"], padding=1), 0.2)

        a[0] = conv(0, in2); a[1] = conv(1, a[0]); a[2] = conv(2, a[1])
        for lv in range(1, 5):
            l0 = 3 * lv
            a[l0] = conv(l0, F.max_pool2d(a[l0 - 1], 2)); a[l0 + 1] = conv(l0 + 1, a[l0]); a[l0 + 2] = conv(l0 + 2, a[l0 + 1])
        for k in range(4):
            lv, l0 = 3 - k, 15 + 3 * k
            up = F.interpolate(a[l0 - 1], scale_factor=2, mode="bilinear", align_corners=True)
            a[l0] = conv(l0, torch.cat([a[3 * lv + 2], up], dim=1)); a[l0 + 1] = conv(l0 + 1, a[l0]); a[l0 + 2] = conv(l0 + 2, a[l0 + 1])
        w_out = sd["outc.conv.weight"].reshape(1, 32, 1, 1)
        r = x + F.conv2d(a[26], sd["outc.conv.weight"], sd["outc.conv.bias"])
        # backward
        gr = torch.where((r >= 0) & (r <= 1), gout, torch.zeros_like(gout))
        cur = w_out * gr * _lrelu_d(a[26])
        gcat = [None] * 4

        def dgrad(l, g):
            wt = sd[keys[l] + ".weight"].transpose(0, 1).flip(2, 3)      # [Cin][Cout] transposed, taps flipped
            return F.conv2d(g, wt, padding=1)

        for l in range(26, 0, -1):
            cin, cout, lv = _SPECS[l]
            if l >= 15 and (l - 15) % 3 == 0:
                gcat[lv] = dgrad(l, cur)
                gup = gcat[lv][:, _CH[lv]:]
                h, w = gup.shape[2] // 2, gup.shape[3] // 2
                My, Mx = _up_matrix(h), _up_matrix(w)
                cur = torch.einsum("Yi,bcYX,Xj->bcij", My, gup, Mx) * _lrelu_d(a[l - 1])
            elif l <= 12 and l % 3 == 0:
                gpool = dgrad(l, cur)
                cur = (_pool_bwd(gpool, a[l - 1]) + gcat[lv - 1][:, :_CH[lv - 1]]) * _lrelu_d(a[l - 1])
            else:
                cur = dgrad(l, cur) * _lrelu_d(a[l - 1])
        gin2 = dgrad(0, cur)
        return gin2[:, :1] + gr, gin2[:, 1].reshape(B, -1).sum(1)


# ----------------------------------------------------------------------------
# SPI (tasks/spi/solver.py:17-51): only the closed-form branch of spi_inverse carries a gradient
# ----------------------------------------------------------------------------

def admm_spi_vjp_autograd(sd, state, x0, K, sigma_d, mu, gout):
    state = state.detach().clone().requires_grad_(True)
    sigma_d = sigma_d.detach().clone().requires_grad_(True)
    mu = mu.detach().clone().requires_grad_(True)
    with torch.enable_grad():
        out = O.admm_spi(sd, state, x0, K, sigma_d, mu)
        gs, gm, gst = torch.autograd.grad(out, (sigma_d, mu, state), gout)
    return gs, gm, gst


def admm_spi_trajectory(sd, state, x0, K, sigma_d, mu):
    states = [state]
    with torch.no_grad():
        for i in range(sigma_d.shape[-1]):
            states.append(O.admm_spi(sd, states[-1], x0, K, sigma_d[:, i:i + 1], mu[:, i:i + 1]))
    return states


def admm_spi_vjp_manual(sd, states, x0, K, sigma_d, mu, gout, denoise_vjp=denoise_vjp_autograd):
    """The recursion of grad_elem.cuh: spi_backward_sequence / spi_step_elem."""
    B, it = sigma_d.shape
    gx, gz, gu = (t.clone() for t in O._split3(gout))
    g_sigma, g_mu = torch.zeros(B, it), torch.zeros(B, it)
    Kv = K[:, 0, 0, 0].reshape(B, 1, 1, 1) * 10
    Ksq = Kv ** 2
    K1 = x0 * Ksq
    for i in reversed(range(it)):
        x, _, u = O._split3(states[i])
        _, zn, un = O._split3(states[i + 1])
        m = mu[:, i].reshape(B, 1, 1, 1)
        gv, gs = denoise_vjp(sd, zn - un, sigma_d[:, i], gx)
        g_sigma[:, i] = gs
        gut = gu - gv
        gzt = gz + gv - gut
        K0 = Ksq - K1
        zpre = (x + u) - K0 / m
        live = (K1 == 0) & (zpre >= 0) & (zpre <= 1)
        gt = torch.where(live, gzt, torch.zeros_like(gzt))
        g_mu[:, i] = (gt * K0 / (m * m)).reshape(B, -1).sum(1)
        gx = gut + gt
        gu = gut + gt
        gz = torch.zeros_like(gz)
    return g_sigma, g_mu, torch.cat((gx, gz, gu), dim=1)


# ----------------------------------------------------------------------------
# CT (tasks/ct/solver.py:17-53) on the build's own Radon pair: parity unpinned like the forward path, so the gradients
# are checked against autograd through the oracle only
# ----------------------------------------------------------------------------

def iadmm_ct_vjp_autograd(sd, state, y0, views, opnorm, sigma_d, mu, tau, gout):
    state = state.detach().clone().requires_grad_(True)
    ps = [p.detach().clone().requires_grad_(True) for p in (sigma_d, mu, tau)]
    with torch.enable_grad():
        out = O.iadmm_ct(sd, state, y0, views, opnorm, *ps)
        gs, gm, gt, gst = torch.autograd.grad(out, (*ps, state), gout)
    return gs, gm, gt, gst


def iadmm_ct_trajectory(sd, state, y0, views, opnorm, sigma_d, mu, tau):
    states = [state]
    with torch.no_grad():
        for i in range(sigma_d.shape[-1]):
            states.append(O.iadmm_ct(sd, states[-1], y0, views, opnorm, sigma_d[:, i:i + 1], mu[:, i:i + 1], tau[:, i:i + 1]))
    return states


# ----------------------------------------------------------------------------
# PR (tasks/pr/solver.py:37-76)
# ----------------------------------------------------------------------------

def iadmm_pr_vjp_autograd(sd, state, y0, mask, sigma_d, mu, tau, gout):
    state = state.detach().clone().requires_grad_(True)
    ps = [p.detach().clone().requires_grad_(True) for p in (sigma_d, mu, tau)]
    with torch.enable_grad():
        out = O.iadmm_pr(sd, state, y0, mask, *ps)
        gs, gm, gt, gst = torch.autograd.grad(out, (*ps, state), gout)
    return gs, gm, gt, gst


def iadmm_pr_trajectory(sd, state, y0, mask, sigma_d, mu, tau):
    states = [state]
    with torch.no_grad():
        for i in range(sigma_d.shape[-1]):
            states.append(O.iadmm_pr(sd, states[-1], y0, mask, sigma_d[:, i:i + 1], mu[:, i:i + 1], tau[:, i:i + 1]))
    return states
