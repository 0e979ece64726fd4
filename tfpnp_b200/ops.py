"""Stand-alone native operators: the CT Radon pair and the PSNR reward metric."""
from __future__ import annotations

import math

import torch

from . import _lib


def _tables(views):
    angles = torch.linspace(0, 179 / 180 * math.pi, views, dtype=torch.float32)   # transforms.py:488
    return torch.cos(angles.double()).float().contiguous(), torch.sin(angles.double()).float().contiguous()


def det_count(resolution: int) -> int:
    return int(math.ceil(math.sqrt(2) * resolution))                              # transforms.py:489


def radon_forward(img: torch.Tensor, views: int) -> torch.Tensor:
    """A: [B,1,N,N] -> [B,1,views,det]  (role of torch_radon's Radon.forward, transforms.py:465-491)."""
    assert img.is_cuda and img.dim() == 4 and img.shape[1] == 1 and img.shape[2] == img.shape[3]
    B, _, N, _ = img.shape
    img = img.contiguous().float()
    out = torch.empty(B, 1, views, det_count(N), device=img.device, dtype=torch.float32)
    cos, sin = _tables(views)
    with torch.cuda.device(img.device):
        _lib.check(_lib.lib().tfpnp_radon_forward(img.data_ptr(), out.data_ptr(), B, N, views, cos.data_ptr(),
                                                  sin.data_ptr(), torch.cuda.current_stream().cuda_stream),
                   "tfpnp_radon_forward")
    return out


def radon_backward(sino: torch.Tensor, resolution: int, views: int) -> torch.Tensor:
    """A^T: [B,1,views,det] -> [B,1,N,N]  (role of Radon.backprojection)."""
    assert sino.is_cuda and sino.dim() == 4 and sino.shape[2] == views and sino.shape[3] == det_count(resolution)
    B = sino.shape[0]
    sino = sino.contiguous().float()
    out = torch.empty(B, 1, resolution, resolution, device=sino.device, dtype=torch.float32)
    cos, sin = _tables(views)
    with torch.cuda.device(sino.device):
        _lib.check(_lib.lib().tfpnp_radon_backward(sino.data_ptr(), out.data_ptr(), B, resolution, views,
                                                   cos.data_ptr(), sin.data_ptr(),
                                                   torch.cuda.current_stream().cuda_stream),
                   "tfpnp_radon_backward")
    return out


def _psnr_native(o: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    N = o.shape[0]
    res = torch.empty(N, device=o.device, dtype=torch.float32)
    with torch.cuda.device(o.device):
        _lib.check(_lib.lib().tfpnp_psnr(o.data_ptr(), g.data_ptr(), res.data_ptr(), N, o.shape[1],
                                         torch.cuda.current_stream().cuda_stream), "tfpnp_psnr")
    return res


class _PsnrFn(torch.autograd.Function):
    """torch_psnr under autograd (the reward enters the actor loss, tfpnp/trainer/mddpg/trainer.py:189): native forward,
    tfpnp_psnr_backward (validated on the GPU in round 2 against autograd through the reference, tests/test_grad.py)."""

    @staticmethod
    def forward(ctx, o, g):
        res = _psnr_native(o, g)
        ctx.save_for_backward(o, g, res)
        return res

    @staticmethod
    def backward(ctx, gres):
        o, g, res = ctx.saved_tensors
        go = torch.empty_like(o)
        gres = gres.contiguous().float()
        with torch.cuda.device(o.device):
            _lib.check(_lib.lib().tfpnp_psnr_backward(o.data_ptr(), g.data_ptr(), res.data_ptr(), gres.data_ptr(), go.data_ptr(),
                                                      o.shape[0], o.shape[1], torch.cuda.current_stream().cuda_stream),
                       "tfpnp_psnr_backward")
        return go, None


def torch_psnr(output: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """Per-image PSNR [B,1], tfpnp/env/base.py:237-242, one fused kernel (differentiable w.r.t. ``output``)."""
    assert output.is_cuda and gt.is_cuda
    N = output.shape[0]
    o = output.contiguous().float().reshape(N, -1)
    g = gt.detach().contiguous().float().reshape(N, -1)
    if torch.is_grad_enabled() and o.requires_grad:
        return _PsnrFn.apply(o, g).unsqueeze(1)
    return _psnr_native(o, g).unsqueeze(1)


def conv3x3_lrelu_nhwc(x0: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, x1: torch.Tensor = None):
    """One UNet ConvLayer (unet.py:8-22) on the tensor cores: x0 [B,H,W,C0] (+ x1 [B,H,W,C1]) fp16 NHWC,
    ``weight`` [Cout, C0+C1, 3, 3] fp32 (re-laid-out to [tap][Cout][Cin] fp16 here) -> [B,H,W,Cout] fp16."""
    assert x0.is_cuda and x0.dtype == torch.float16 and x0.is_contiguous()
    B, H, W, C0 = x0.shape
    C1 = 0 if x1 is None else x1.shape[-1]
    Cout = weight.shape[0]
    assert weight.shape[1] == C0 + C1
    wt = weight.permute(2, 3, 0, 1).reshape(9, Cout, C0 + C1).to(torch.float16).contiguous().to(x0.device)
    b = bias.float().contiguous().to(x0.device)
    out = torch.empty(B, H, W, Cout, device=x0.device, dtype=torch.float16)
    with torch.cuda.device(x0.device):
        _lib.check(_lib.lib().tfpnp_conv3x3_nhwc(x0.data_ptr(), C0, x1.data_ptr() if x1 is not None else None, C1,
                                                 wt.data_ptr(), b.data_ptr(), out.data_ptr(), B, H, W, Cout,
                                                 torch.cuda.current_stream().cuda_stream), "tfpnp_conv3x3_nhwc")
    return out


# ---- stand-alone transforms (tfpnp/utils/transforms.py) ----------------------------------------------------

def _fft2_native(x: torch.Tensor, inverse: bool, centered: bool) -> torch.Tensor:
    if not x.is_cuda:
        raise RuntimeError("tfpnp_b200 transforms run on CUDA (sm_100) tensors only; there is no CPU fallback")
    assert x.shape[-1] == 2 and x.shape[-2] == x.shape[-3], "expected [..., N, N, 2]"
    N = x.shape[-2]
    xin = x.contiguous().float()
    n = xin.numel() // (N * N * 2)
    out = torch.empty_like(xin)
    ws = torch.empty_like(xin)
    if n:
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().tfpnp_fft2(xin.data_ptr(), out.data_ptr(), ws.data_ptr(), n, N, 1 if inverse else 0,
                                             1 if centered else 0, torch.cuda.current_stream().cuda_stream), "tfpnp_fft2")
    return out


def fft2(data: torch.Tensor) -> torch.Tensor:
    """transforms.fft2 (transforms.py:68-84): centred ortho 2-D FFT of [..., N, N, 2]."""
    return _fft2_native(data, False, True)


def ifft2(data: torch.Tensor) -> torch.Tensor:
    """transforms.ifft2 (transforms.py:87-103)."""
    return _fft2_native(data, True, True)


def complex_mul(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """transforms.complex_mul (transforms.py:260-270)."""
    re = x[..., 0] * y[..., 0] - x[..., 1] * y[..., 1]
    im = x[..., 0] * y[..., 1] + x[..., 1] * y[..., 0]
    return torch.stack((re, im), dim=-1)


def cdp_forward(data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """transforms.cdp_forward (transforms.py:282-301): FFT2_ortho(data * mask_j), j = 1..M, un-centred.
    data [B,1,N,N,2], mask [B,M,N,N,2] -> [B,M,N,N,2]."""
    if data.dim() == 4:
        data = torch.stack([data, torch.zeros_like(data)], -1)
    return _fft2_native(complex_mul(data, mask), False, False)


def cdp_backward(data: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """transforms.cdp_backward (transforms.py:304-320): mean_j(IFFT2_ortho(data_j) * conj(mask_j)) -> [B,1,N,N,2]."""
    conj = torch.stack((mask[..., 0], -mask[..., 1]), dim=-1)
    return complex_mul(_fft2_native(data, True, False), conj).mean(dim=1, keepdim=True)
