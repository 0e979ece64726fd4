"""CPU restatement of the TFPnP PnP-ADMM hot path.  TEST INFRASTRUCTURE ONLY.

Plain PyTorch-on-CPU (fp32 by default, fp64 on request for error budgeting)
restatement of the reference algorithm, so that it can travel to the GPU box
where ``/root/reference`` does not exist.  Every function cites the reference
file:line it follows (paths relative to the reference checkout).

Pinning status
--------------
* CS-MRI / PR / SPI solvers, the UNet denoiser and PSNR are pinned against the
  UNMODIFIED reference classes executed through ``oracle/refshim.py`` in the
  build container: ``oracle/make_golden.py`` stores the reference outputs under
  ``tests/golden/`` and ``tests/test_oracle_golden.py`` checks this file against
  them.  (The reference ships no tests / golden vectors of its own)
* CT: **parity unpinned**.  The arithmetic of the reference's CT path lives in
  the third-party package ``torch_radon`` (unpinned, absent from the reference
  checkout and from this image).  ``radon_forward`` / ``radon_backward`` restate
  the *geometry* the reference fixes (tfpnp/utils/transforms.py:487-491) with a
  Joseph-type discretisation chosen by this build and validated structurally
  (adjointness, analytic disk sinogram).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# ----------------------------------------------------------------------------
# UNet(2,1) denoiser  (tfpnp/pnp/denoiser/models/unet.py:8-131,
#                      tfpnp/pnp/denoiser/base.py:23-32)
# ----------------------------------------------------------------------------

#: (block name, in_ch, out_ch) in state_dict order (unet.py:37-46)
UNET_BLOCKS = (
    ("inc.conv", 2, 32),
    ("down1.mpconv.1", 32, 64),
    ("down2.mpconv.1", 64, 128),
    ("down3.mpconv.1", 128, 256),
    ("down4.mpconv.1", 256, 512),
    ("up1.conv", 768, 256),
    ("up2.conv", 384, 128),
    ("up3.conv", 192, 64),
    ("up4.conv", 96, 32),
)


def unet_param_shapes():
    """Ordered (key, shape) list of UNet(2,1).state_dict() (unet.py:34-47)."""
    out = []
    for name, cin, cout in UNET_BLOCKS:
        for k in range(3):
            ci = cin if k == 0 else cout
            out.append((f"{name}.conv-{k}.conv2d.weight", (cout, ci, 3, 3)))
            out.append((f"{name}.conv-{k}.conv2d.bias", (cout,)))
    out.append(("outc.conv.weight", (1, 32, 1, 1)))
    out.append(("outc.conv.bias", (1,)))
    return out


def _conv_block(sd, name, x, quant):
    # ConvBlock = 3 x [conv3x3 pad 1 + bias, LeakyReLU(0.2)]  (unet.py:20-31)
    for k in range(3):
        w = sd[f"{name}.conv-{k}.conv2d.weight"]
        b = sd[f"{name}.conv-{k}.conv2d.bias"]
        if quant is not None:
            x, w = quant(x), quant(w)
        x = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.2)
    return x


def unet_forward(sd: Dict[str, Tensor], x: Tensor,
                 quant: Optional[Callable[[Tensor], Tensor]] = None) -> Tensor:
    """UNet.forward (unet.py:52-66).  ``quant`` optionally rounds conv operands
    (used by tests to predict the error of reduced-precision tensor-core modes)."""
    noisy = x
    x1 = _conv_block(sd, "inc.conv", x, quant)
    skips = [x1]
    h = x1
    for d in ("down1", "down2", "down3", "down4"):        # unet.py:80-90
        h = _conv_block(sd, f"{d}.mpconv.1", F.max_pool2d(h, 2), quant)
        skips.append(h)
    h = skips.pop()
    for u in ("up1", "up2", "up3", "up4"):                # unet.py:93-121
        skip = skips.pop()
        h = F.interpolate(h, scale_factor=2, mode="bilinear", align_corners=True)
        dy, dx = skip.shape[2] - h.shape[2], skip.shape[3] - h.shape[3]
        if dy or dx:
            h = F.pad(h, (dx // 2, dx - dx // 2, dy // 2, dy - dy // 2))
        h = _conv_block(sd, f"{u}.conv", torch.cat([skip, h], dim=1), quant)
    w, b = sd["outc.conv.weight"], sd["outc.conv.bias"]
    if quant is not None:
        h, w = quant(h), quant(w)
    residual = F.conv2d(h, w, b)                            # unet.py:124-131
    return noisy[:, :1] + residual                          # unet.py:65-66


def denoise(sd: Dict[str, Tensor], x: Tensor, sigma: Tensor, quant=None) -> Tensor:
    """UNetDenoiser2D.forward (denoiser/base.py:23-32): x [B,1,H,W], sigma [B]."""
    if "model.0.weight" in sd:          # an IRCNN state_dict (self-defined, see ircnn_denoise)
        return ircnn_denoise(sd, x, sigma)
    n, _, h, w = x.shape
    noise_map = torch.ones(n, 1, h, w, dtype=x.dtype) * sigma.reshape(n, 1, 1, 1).to(x.dtype)
    out = unet_forward(sd, torch.cat([x, noise_map], dim=1), quant)
    return torch.clamp(out, 0, 1)


# ----------------------------------------------------------------------------
# transforms  (tfpnp/utils/transforms.py)
# ----------------------------------------------------------------------------

def _c(x: Tensor) -> Tensor:
    return torch.view_as_complex(x.contiguous())


def _r(x: Tensor) -> Tensor:
    return torch.view_as_real(x)


def fft2c(x: Tensor) -> Tensor:
    """Centred ortho 2-D FFT over dims (-3,-2) of a [...,H,W,2] tensor
    (transforms.py:68-84; shifts :215-257)."""
    c = torch.fft.ifftshift(_c(x), dim=(-2, -1))
    c = torch.fft.fft2(c, norm="ortho")
    return _r(torch.fft.fftshift(c, dim=(-2, -1)))


def ifft2c(x: Tensor) -> Tensor:
    """Centred ortho inverse 2-D FFT (transforms.py:87-103)."""
    c = torch.fft.ifftshift(_c(x), dim=(-2, -1))
    c = torch.fft.ifft2(c, norm="ortho")
    return _r(torch.fft.fftshift(c, dim=(-2, -1)))


def real2complex(x: Tensor) -> Tensor:          # transforms.py:12-13
    return torch.stack([x, torch.zeros_like(x)], dim=-1)


def complex2real(x: Tensor) -> Tensor:          # transforms.py:16-17
    return x[..., 0]


def _cmul(a: Tensor, b: Tensor) -> Tensor:      # transforms.py:260-270
    return torch.stack((a[..., 0] * b[..., 0] - a[..., 1] * b[..., 1],
                        a[..., 0] * b[..., 1] + a[..., 1] * b[..., 0]), -1)


def cdp_forward(z: Tensor, mask: Tensor) -> Tensor:
    """FFT2_ortho(z * mask_j), un-centred (transforms.py:282-301)."""
    zz = z.expand(-1, mask.shape[1], -1, -1, -1)
    return _r(torch.fft.fft2(_c(_cmul(zz, mask)), norm="ortho"))


def cdp_backward(g: Tensor, mask: Tensor) -> Tensor:
    """mean_j(IFFT2_ortho(g_j) * conj(mask_j)) (transforms.py:304-320)."""
    t = _r(torch.fft.ifft2(_c(g), norm="ortho"))
    conj = torch.stack([mask[..., 0], -mask[..., 1]], -1)
    return _cmul(t, conj).mean(1, keepdim=True)


def spi_inverse(ztilde: Tensor, K1: Tensor, K: Tensor, mu: Tensor) -> Tensor:
    """Prox of the quanta-image-sensor likelihood (transforms.py:404-439).

    Per pixel: closed form where K1 == 0, otherwise 10 bisection steps on
    [1e-5, 1.1] for the root of f(y) = K1/(e^y - 1) - mu*y - K0 + mu*ztilde."""
    dt = ztilde.dtype
    K0 = K ** 2 - K1
    closed = ztilde - K0 / mu
    frozen = (K1 == 0).expand_as(ztilde).clone()
    bmin = torch.full_like(ztilde, 1e-5)
    bmax = torch.full_like(ztilde, 1.1)
    bave = (bmin + bmax) / 2.0
    for _ in range(10):
        f = K1 / (torch.exp(bave) - 1) - mu * bave - K0 + mu * ztilde
        live = ~frozen
        pos, neg = (f > 0) & live, (f < 0) & live
        frozen = frozen | ((f == 0) & live)
        bmin = torch.where(pos, bave, bmin)
        bmax = torch.where(neg, bave, bmax)
        bave = torch.where(~frozen, (bmin + bmax) / 2.0, bave)
    z = torch.where((K1 != 0).expand_as(ztilde), bave, closed)
    return torch.clamp(z, 0.0, 1.0).to(dt)


# ----------------------------------------------------------------------------
# CT: parallel-beam Radon pair (geometry: transforms.py:487-491).
# PARITY UNPINNED -- torch_radon is absent; discretisation chosen by this build.
# ----------------------------------------------------------------------------

def ct_geometry(resolution: int, views: int):
    """angles = linspace(0, 179*pi/180, views); det_count = ceil(sqrt(2)*res);
    unit detector spacing (transforms.py:487-491).  Returns fp32 cos/sin tables
    (computed in fp64 from the fp32 angles) and det_count."""
    angles = torch.linspace(0, 179 / 180 * math.pi, views, dtype=torch.float32)
    det = int(math.ceil(math.sqrt(2) * resolution))
    cs = torch.cos(angles.double()).float()
    sn = torch.sin(angles.double()).float()
    return cs, sn, det


def radon_forward(img: Tensor, cs: Tensor, sn: Tensor, det: int) -> Tensor:
    """A: [B,1,N,N] -> [B,1,V,D].  Joseph-type ray-driven projector: for each ray
    x*cos + y*sin = s, step over the pixels of the driving axis (columns when
    |sin| >= |cos|, rows otherwise), interpolate linearly between the two
    nearest pixels of the other axis, weight by 1/max(|cos|,|sin|).  Pixel (i,j)
    sits at x = j - c, y = i - c, c = (N-1)/2; detector d at s = d - (D-1)/2."""
    B, _, N, _ = img.shape
    dt = img.dtype
    c = (N - 1) / 2.0
    s = torch.arange(det, dtype=dt) - (det - 1) / 2.0        # [D]
    t = torch.arange(N, dtype=dt) - c                         # driving coordinate
    out = torch.zeros(B, 1, len(cs), det, dtype=dt)
    flat = img.reshape(B, N, N)
    for v in range(len(cs)):
        co, si = cs[v].to(dt), sn[v].to(dt)
        col_drive = bool(abs(float(si)) >= abs(float(co)))
        m = abs(si) if col_drive else abs(co)
        # continuous index along the interpolated axis, [D, N]
        if col_drive:
            r = (s[:, None] - t[None, :] * co) / si + c       # row index at column j
        else:
            r = (s[:, None] - t[None, :] * si) / co + c       # col index at row i
        i0 = torch.floor(r)
        f = r - i0
        i0 = i0.long()
        acc = torch.zeros(B, det, dtype=dt)
        for k, wgt in ((0, 1 - f), (1, f)):
            idx = i0 + k
            ok = (idx >= 0) & (idx < N)
            idc = idx.clamp(0, N - 1)
            drv = torch.arange(N).expand(det, N)
            if col_drive:
                vals = flat[:, idc, drv]                      # [B, D, N]
            else:
                vals = flat[:, drv, idc]
            acc = acc + (vals * (wgt * ok)[None]).sum(-1)
        out[:, 0, v] = acc / m
    return out


def radon_backward(sino: Tensor, cs: Tensor, sn: Tensor, N: int) -> Tensor:
    """A^T: [B,1,V,D] -> [B,1,N,N]; the exact transpose of ``radon_forward``
    in gather form: pixel (i,j) projects to d* = x*cos + y*sin + (D-1)/2 and
    receives hat((d - d*)/m)/m * sino[v,d] from d in {floor(d*), floor(d*)+1},
    m = max(|cos|,|sin|)."""
    B, _, V, det = sino.shape
    dt = sino.dtype
    c = (N - 1) / 2.0
    yy = (torch.arange(N, dtype=dt) - c)[:, None].expand(N, N)
    xx = (torch.arange(N, dtype=dt) - c)[None, :].expand(N, N)
    out = torch.zeros(B, N, N, dtype=dt)
    for v in range(V):
        co, si = cs[v].to(dt), sn[v].to(dt)
        m = max(abs(si), abs(co))
        dstar = xx * co + yy * si + (det - 1) / 2.0
        d0 = torch.floor(dstar)
        for k in (0, 1):
            d = d0 + k
            wgt = torch.clamp(1 - (d - dstar).abs() / m, min=0) / m
            ok = (d >= 0) & (d < det)
            di = d.long().clamp(0, det - 1)
            out = out + sino[:, 0, v][:, di] * (wgt * ok)[None]
    return out.reshape(B, 1, N, N)


def radon_opnorm(N: int, cs: Tensor, sn: Tensor, det: int, seed: int = 0,
                 n_iter: int = 10) -> float:
    """sqrt(lambda_max(A^T A)) by the reference's power method
    (transforms.py:447-462) from a SEEDED start vector (the reference uses an
    unseeded torch.randn on the GPU, :468-472)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(1, 1, N, N, generator=g)
    x = x / x.norm()
    v = 0.0
    for _ in range(n_iter):
        x = radon_backward(radon_forward(x, cs, sn, det), cs, sn, N)
        v = float(x.norm())
        x = x / v
    return v ** 0.5


# ----------------------------------------------------------------------------
# IRCNN.  ABSENT from the reference (tfpnp/pnp/__init__.py:5-13): PARITY UNPINNED.
# Restates the published network (Zhang et al., CVPR 2017; inference form, BatchNorm folded): seven
# 3x3 convolutions, 64 channels, dilations 1,2,3,4,3,2,1, ReLU between them, residual output, wrapped
# like UNetDenoiser2D (tfpnp/pnp/denoiser/base.py:23-32).
# ----------------------------------------------------------------------------
IRCNN_DILATIONS = (1, 2, 3, 4, 3, 2, 1)


def ircnn_param_shapes():
    out = []
    for i, (ci, co) in enumerate([(2, 64)] + [(64, 64)] * 5 + [(64, 1)]):
        out += [(f"model.{2 * i}.weight", (co, ci, 3, 3)), (f"model.{2 * i}.bias", (co,))]
    return out


def ircnn_denoise(sd: Dict[str, Tensor], x: Tensor, sigma: Tensor) -> Tensor:
    N, _, H, W = x.shape
    h = torch.cat([x, torch.ones(N, 1, H, W, dtype=x.dtype) * sigma.reshape(N, 1, 1, 1).to(x.dtype)], dim=1)
    for i, d in enumerate(IRCNN_DILATIONS):
        h = F.conv2d(h, sd[f"model.{2 * i}.weight"].to(x.dtype), sd[f"model.{2 * i}.bias"].to(x.dtype), padding=d, dilation=d)
        if i < 6:
            h = F.relu(h)
    return torch.clamp(x - h, 0, 1)


# ----------------------------------------------------------------------------
# solver state helpers (tfpnp/pnp/solver/base.py:87-116)
# ----------------------------------------------------------------------------

def admm_reset(x0: Tensor) -> Tensor:           # base.py:95-99
    x = x0.clone()
    return torch.cat((x, x.clone(), torch.zeros_like(x)), dim=1)


def pr_reset(x0: Tensor) -> Tensor:             # tasks/pr/solver.py:29-35
    return admm_reset(real2complex(x0))


def _split3(state: Tensor):
    n = state.shape[1] // 3
    return torch.split(state, n, dim=1)


def get_output(state: Tensor, complex_state: bool) -> Tensor:
    """ADMMSolver.get_output (base.py:101-104), with the real part taken for the
    complex tasks (tasks/csmri/solver.py:9-18, tasks/pr/solver.py:9-19)."""
    x = _split3(state)[0]
    return complex2real(x) if complex_state else x


# ----------------------------------------------------------------------------
# the four inner loops
# ----------------------------------------------------------------------------

def admm_csmri(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, mu: Tensor,
               iter_num: Optional[int] = None, quant=None) -> Tensor:
    """ADMMSolver_CSMRI.forward (tasks/csmri/solver.py:29-57)."""
    x, z, u = _split3(state)
    B = x.shape[0]
    it = sigma_d.shape[-1] if iter_num is None else iter_num
    m = mask.bool()[..., None].expand_as(y0)
    for i in range(it):
        x = real2complex(denoise(sd, complex2real(z - u), sigma_d[:, i], quant))
        Z = fft2c(x + u)
        _mu = mu[:, i].reshape(B, 1, 1, 1, 1).to(Z.dtype)
        Z = torch.where(m, (_mu * Z + y0) / (1 + _mu), Z)
        z = ifft2c(Z)
        u = u + x - z
    return torch.cat((x, z, u), dim=1)


def _dc_blend(Z: Tensor, y0: Tensor, m: Tensor, mu: Tensor) -> Tensor:
    return torch.where(m, (mu * Z + y0) / (1 + mu), Z)


def hqs_csmri(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, mu: Tensor) -> Tensor:
    """HQSSolver_CSMRI.forward (tasks/csmri/solver.py:60-88); state = cat[x, z]."""
    x, z = torch.split(state, state.shape[1] // 2, dim=1)
    B = x.shape[0]
    m = mask.bool()[..., None].expand_as(y0)
    for i in range(sigma_d.shape[-1]):
        x = real2complex(denoise(sd, complex2real(z), sigma_d[:, i]))
        z = ifft2c(_dc_blend(fft2c(x), y0, m, mu[:, i].reshape(B, 1, 1, 1, 1)))
    return torch.cat([x, z], dim=1)


def pg_csmri(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, tau: Tensor) -> Tensor:
    """PGSolver_CSMRI.forward (tasks/csmri/solver.py:91-118); state = x."""
    x = state
    B = x.shape[0]
    m = mask.bool()[..., None].expand_as(y0)
    for i in range(sigma_d.shape[-1]):
        temp = torch.where(m, fft2c(x) - y0, torch.zeros_like(y0))
        z = x - tau[:, i].reshape(B, 1, 1, 1, 1) * ifft2c(temp)
        x = real2complex(denoise(sd, complex2real(z), sigma_d[:, i]))
    return x


def apg_csmri(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, tau: Tensor, beta: Tensor) -> Tensor:
    """APGSolver_CSMRI.forward (tasks/csmri/solver.py:121-161); state = cat[x, s]."""
    x, s = torch.split(state, state.shape[1] // 2, dim=1)
    B = x.shape[0]
    m = mask.bool()[..., None].expand_as(y0)
    for i in range(sigma_d.shape[-1]):
        temp = torch.where(m, fft2c(s) - y0, torch.zeros_like(y0))
        z = s - tau[:, i].reshape(B, 1, 1, 1, 1) * ifft2c(temp)
        x_prev = x
        x = real2complex(denoise(sd, complex2real(z), sigma_d[:, i]))
        s = x + beta[:, i].reshape(B, 1, 1, 1, 1) * (x - x_prev)
    return torch.cat([x, s], dim=1)


def redadmm_csmri(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, mu: Tensor, lamda: Tensor) -> Tensor:
    """REDADMMSolver_CSMRI.forward (tasks/csmri/solver.py:164-201); state = cat[x, z, u]."""
    x, z, u = _split3(state)
    B = x.shape[0]
    m = mask.bool()[..., None].expand_as(y0)
    for i in range(sigma_d.shape[-1]):
        _mu = mu[:, i].reshape(B, 1, 1, 1, 1)
        _l = lamda[:, i].reshape(B, 1, 1, 1, 1)
        x_half = real2complex(denoise(sd, complex2real(x), sigma_d[:, i]))
        x = (_l * x_half + _mu * (z - u)) / (_mu + _l)
        z = ifft2c(_dc_blend(fft2c(x + u), y0, m, _mu))
        u = u + x - z
    return torch.cat([x, z, u], dim=1)


def iadmm_pr(sd, state: Tensor, y0: Tensor, mask: Tensor, sigma_d: Tensor, mu: Tensor,
             tau: Tensor, iter_num: Optional[int] = None, quant=None) -> Tensor:
    """IADMMSolver_PR.forward (tasks/pr/solver.py:37-76)."""
    x, z, u = _split3(state)
    B = x.shape[0]
    it = sigma_d.shape[-1] if iter_num is None else iter_num
    for i in range(it):
        x = real2complex(denoise(sd, complex2real(z - u), sigma_d[:, i], quant))
        _tau = tau[:, i].reshape(B, 1, 1, 1, 1).to(z.dtype)
        _mu = mu[:, i].reshape(B, 1, 1, 1, 1).to(z.dtype)
        Az = cdp_forward(z, mask)
        y_hat = (Az ** 2).sum(dim=-1).sqrt()                 # transforms.py:106-118
        ratio = (y_hat - y0) / y_hat                         # unguarded /0, solver.py:67
        g = cdp_backward(torch.stack((ratio * Az[..., 0], ratio * Az[..., 1]), -1), mask)
        z = z - _tau * (g + _mu * (z - (x + u)))
        u = u + x - z
    return torch.cat((x, z, u), dim=1)


def iadmm_ct(sd, state: Tensor, y0: Tensor, views: int, opnorm: float, sigma_d: Tensor,
             mu: Tensor, tau: Tensor, iter_num: Optional[int] = None, quant=None) -> Tensor:
    """IADMMSolver_CT.forward (tasks/ct/solver.py:17-53) with the build's own
    Radon pair and an explicit ``opnorm`` (see module docstring)."""
    x, z, u = _split3(state)
    B, _, N, _ = x.shape
    cs, sn, det = ct_geometry(N, views)
    it = sigma_d.shape[-1] if iter_num is None else iter_num
    for i in range(it):
        x = denoise(sd, z - u, sigma_d[:, i], quant)
        _tau = tau[:, i].reshape(B, 1, 1, 1).to(z.dtype)
        _mu = mu[:, i].reshape(B, 1, 1, 1).to(z.dtype)
        bp = radon_backward(radon_forward(z, cs, sn, det) - y0, cs, sn, N) / opnorm ** 2
        z = z - _tau * (bp + _mu * (z - (x + u)))
        u = u + x - z
    return torch.cat((x, z, u), dim=1)


def pg_ct(sd, state: Tensor, y0: Tensor, views: int, opnorm: float, sigma_d: Tensor, tau: Tensor) -> Tensor:
    """PGSolver_CT.forward (tasks/ct/solver.py:56-87) on this build's Radon pair (PARITY UNPINNED w.r.t. torch_radon)."""
    x = state
    B, n = x.shape[0], x.shape[-1]
    cs, sn, det = ct_geometry(n, views)
    for i in range(sigma_d.shape[-1]):
        g = radon_backward(radon_forward(x, cs, sn, det) - y0, cs, sn, n) / opnorm ** 2
        z = x - tau[:, i].reshape(B, 1, 1, 1) * g
        x = denoise(sd, z, sigma_d[:, i])
    return x


def admm_spi(sd, state: Tensor, x0: Tensor, K: Tensor, sigma_d: Tensor, mu: Tensor,
             iter_num: Optional[int] = None, quant=None) -> Tensor:
    """ADMMSolver_SPI.forward (tasks/spi/solver.py:17-51); order z, u, then x."""
    x, z, u = _split3(state)
    B = x.shape[0]
    it = sigma_d.shape[-1] if iter_num is None else iter_num
    Kv = K[:, 0, 0, 0].reshape(B, 1, 1, 1) * 10
    K1 = x0 * (Kv ** 2)
    for i in range(it):
        _mu = mu[:, i].reshape(B, 1, 1, 1).to(x.dtype)
        z = spi_inverse(x + u, K1, Kv, _mu)
        u = u + x - z
        x = denoise(sd, z - u, sigma_d[:, i], quant)
    return torch.cat((x, z, u), dim=1)


def psnr(output: Tensor, gt: Tensor) -> Tensor:
    """torch_psnr (tfpnp/env/base.py:237-242) -> [B,1]."""
    n = output.shape[0]
    o = torch.clamp(output, 0, 1).reshape(n, -1)
    mse = ((o - gt.reshape(n, -1)) ** 2).mean(dim=1)
    return (10 * torch.log10(1.0 / mse)).unsqueeze(1)
