"""B200-native stand-in for ``tfpnp.pnp.denoiser.UNetDenoiser2D``
(reference: tfpnp/pnp/denoiser/base.py:7-32, models/unet.py:34-131).

Same constructor / call signature: ``UNetDenoiser2D(ckpt_path)`` loads a
``UNet(2,1).state_dict()`` checkpoint, ``forward(x, sigma)`` maps ``x [B,1,H,W]``,
``sigma [B]`` to ``clamp(UNet(cat[x, sigma map]), 0, 1)``.  The network itself runs
in libtfpnp_b200 (tcgen05 implicit-GEMM convolutions); the weights are uploaded and
re-laid-out once per device, not broadcast per call.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict

import torch

from . import _lib

#: state_dict order of UNet(2,1) (unet.py:37-46): (block, cin, cout)
_BLOCKS = (("inc.conv", 2, 32), ("down1.mpconv.1", 32, 64), ("down2.mpconv.1", 64, 128),
           ("down3.mpconv.1", 128, 256), ("down4.mpconv.1", 256, 512), ("up1.conv", 768, 256),
           ("up2.conv", 384, 128), ("up3.conv", 192, 64), ("up4.conv", 96, 32))


def unet_state_dict_layout():
    out = []
    for name, cin, cout in _BLOCKS:
        for k in range(3):
            ci = cin if k == 0 else cout
            out.append((f"{name}.conv-{k}.conv2d.weight", (cout, ci, 3, 3)))
            out.append((f"{name}.conv-{k}.conv2d.bias", (cout,)))
    out.append(("outc.conv.weight", (1, 32, 1, 1)))
    out.append(("outc.conv.bias", (1,)))
    return out


def flatten_state_dict(sd) -> torch.Tensor:
    """Concatenate the 56 tensors in state_dict order; validates names and shapes."""
    parts = []
    for key, shape in unet_state_dict_layout():
        if key not in sd:
            raise KeyError(f"UNet(2,1) checkpoint is missing '{key}'")
        t = sd[key]
        if tuple(t.shape) != shape:
            raise ValueError(f"'{key}' has shape {tuple(t.shape)}, expected {shape}")
        parts.append(t.detach().to(torch.float32).cpu().contiguous().reshape(-1))
    return torch.cat(parts).contiguous()


class UNetDenoiser2D(torch.nn.Module):
    def __init__(self, ckpt_path=None, state_dict=None, precision="fp16x3"):
        super().__init__()
        if state_dict is None:
            if ckpt_path is None:
                # the reference falls back to its bundled pretrained file and raises if absent
                # (denoiser/base.py:10-13); this build ships no weights
                raise ValueError('Default ckpt not found, you have to provide a ckpt path')
            state_dict = torch.load(ckpt_path, map_location="cpu")
        if precision not in _lib.PRECISIONS:
            raise ValueError(f"precision must be one of {list(_lib.PRECISIONS)}")
        self.precision = precision
        self._flat = flatten_state_dict(state_dict)
        self._handles = {}   # device index -> c_void_p

    def _handle(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with torch.cuda.device(idx):
                out = C.c_void_p()
                _lib.check(_lib.lib().tfpnp_denoiser_create(self._flat.data_ptr(), self._flat.numel(),
                                                            _lib.PRECISIONS[self.precision], C.byref(out)),
                           "tfpnp_denoiser_create")
            self._handles[idx] = h = out
        return h

    # Reverse mode: on by default, like autograd through the reference module; ``differentiable = False``
    # turns a request for gradients into a NotImplementedError.  It runs on a second, fp32 engine built from the same
    # weights (tfpnp_denoiser_vjp).
    differentiable = True

    def _grad_handle(self, device: torch.device):
        """The fp32 engine that implements tfpnp_denoiser_vjp (this engine itself when precision == 'fp32_simt')."""
        if self.precision == "fp32_simt":
            return self._handle(device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        key = ("grad", idx)
        h = self._handles.get(key)
        if h is None:
            with torch.cuda.device(idx):
                out = C.c_void_p()
                _lib.check(_lib.lib().tfpnp_denoiser_create(self._flat.data_ptr(), self._flat.numel(),
                                                            _lib.PREC_FP32_SIMT, C.byref(out)), "tfpnp_denoiser_create")
            self._handles[key] = h = out
        return h

    def vjp(self, x, sigma, gout):
        """(d<out,gout>/dx [N,1,H,W], d<out,gout>/dsigma [N]) of ``forward`` (denoiser/base.py:23-32 under autograd)."""
        if not x.is_cuda:
            raise RuntimeError("tfpnp_b200.UNetDenoiser2D runs on CUDA (sm_100) tensors only; no CPU fallback")
        N, Cc, H, W = x.shape
        assert Cc == 1
        x = x.detach().contiguous().float()
        sigma = sigma.detach().reshape(N).float().contiguous()
        gout = gout.detach().contiguous().float()
        gx = torch.empty_like(x)
        gs = torch.empty(N, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().tfpnp_denoiser_vjp(self._grad_handle(x.device), x.data_ptr(), sigma.data_ptr(), 1,
                                                     gout.data_ptr(), gx.data_ptr(), gs.data_ptr(), N, H, W, st),
                       "tfpnp_denoiser_vjp")
        return gx, gs

    def forward(self, x, sigma):
        if not x.is_cuda:
            raise RuntimeError("tfpnp_b200.UNetDenoiser2D runs on CUDA (sm_100) tensors only; no CPU fallback")
        if torch.is_grad_enabled() and (x.requires_grad or sigma.requires_grad):
            if not self.differentiable:
                raise NotImplementedError("the differentiable denoiser was switched off (.differentiable = False)")
            return _DenoiseFn.apply(self, x, sigma)
        N, Cc, H, W = x.shape
        assert Cc == 1
        x = x.contiguous().float()
        sigma = sigma.reshape(N).float()
        out = torch.empty_like(x)
        with torch.cuda.device(x.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().tfpnp_denoiser_forward(self._handle(x.device), x.data_ptr(), sigma.data_ptr(),
                                                         sigma.stride(0), out.data_ptr(), N, H, W, st),
                       "tfpnp_denoiser_forward")
        return out

    def __del__(self):
        try:
            for h in self._handles.values():
                _lib.lib().tfpnp_denoiser_destroy(h)
        except Exception:
            pass


class _DenoiseFn(torch.autograd.Function):
    """autograd node for UNetDenoiser2D.forward: the native forward, tfpnp_denoiser_vjp backward."""

    @staticmethod
    def forward(ctx, den, x, sigma):
        with torch.no_grad():
            out = den.forward(x.detach(), sigma.detach())
        ctx.den = den
        ctx.sigma_shape = sigma.shape
        ctx.save_for_backward(x.detach(), sigma.detach())
        return out

    @staticmethod
    def backward(ctx, gout):
        x, sigma = ctx.saved_tensors
        gx, gs = ctx.den.vjp(x, sigma, gout)
        return None, gx.to(x.dtype), gs.reshape(ctx.sigma_shape).to(sigma.dtype)


def random_unet_state_dict(seed: int = 0):
    """A seeded UNet(2,1) state_dict with nn.Conv2d's default initialisation (weights and biases
    U(-1/sqrt(fan_in), 1/sqrt(fan_in))): benchmarks and smoke tests need weights of the right architecture, the
    pretrained unet-nm.pt is not distributed with the reference."""
    import math
    from collections import OrderedDict
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    fan_in = 1
    for key, shape in unet_state_dict_layout():
        if key.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
        b = 1.0 / math.sqrt(fan_in)
        sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * b
    return sd


#: IRCNN(in_nc=2, out_nc=1, nc=64) inference-form state_dict (BatchNorm folded): seven Conv2d at sequential
#: indices 0,2,...,12 with ReLU between them, dilations 1,2,3,4,3,2,1
IRCNN_DILATIONS = (1, 2, 3, 4, 3, 2, 1)


def ircnn_state_dict_layout():
    out = []
    for i, (ci, co) in enumerate([(2, 64)] + [(64, 64)] * 5 + [(64, 1)]):
        out.append((f"model.{2 * i}.weight", (co, ci, 3, 3)))
        out.append((f"model.{2 * i}.bias", (co,)))
    return out


class IRCNNDenoiser2D(UNetDenoiser2D):
    """IRCNN prox_sigma denoiser (BASELINE configs[0]).  The reference has no IRCNN
    (tfpnp/pnp/__init__.py:5-13); this is the published network wrapped like UNetDenoiser2D
    (tfpnp/pnp/denoiser/base.py:23-32): ``clamp(x - net(cat[x, sigma map]), 0, 1)``.  Same call signature,
    usable with every solver of this package."""

    def __init__(self, ckpt_path=None, state_dict=None, precision="fp16x3"):
        torch.nn.Module.__init__(self)
        if state_dict is None:
            if ckpt_path is None:
                raise ValueError('Default ckpt not found, you have to provide a ckpt path')
            state_dict = torch.load(ckpt_path, map_location="cpu")
        if precision not in ("fp16", "fp16x3"):
            raise ValueError("IRCNNDenoiser2D precision must be 'fp16' or 'fp16x3'")
        self.precision = precision
        parts = []
        for key, shape in ircnn_state_dict_layout():
            if key not in state_dict:
                raise KeyError(f"IRCNN checkpoint is missing '{key}'")
            t = state_dict[key]
            if tuple(t.shape) != shape:
                raise ValueError(f"'{key}' has shape {tuple(t.shape)}, expected {shape}")
            parts.append(t.detach().to(torch.float32).cpu().contiguous().reshape(-1))
        self._flat = torch.cat(parts).contiguous()
        self._handles = {}

    def _grad_handle(self, device: torch.device):
        raise NotImplementedError("reverse mode (SURVEY 8f N4) is built for the UNet denoiser only")

    def _handle(self, device: torch.device):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        h = self._handles.get(idx)
        if h is None:
            with torch.cuda.device(idx):
                out = C.c_void_p()
                _lib.check(_lib.lib().tfpnp_ircnn_create(self._flat.data_ptr(), self._flat.numel(),
                                                         _lib.PRECISIONS[self.precision], C.byref(out)),
                           "tfpnp_ircnn_create")
            self._handles[idx] = h = out
        return h


def create_denoiser(opt, ckpt_path=None, state_dict=None, precision="fp16x3"):
    """Mirror of tfpnp.pnp.create_denoiser (tfpnp/pnp/__init__.py:5-13), plus 'ircnn'."""
    print(f'[i] use denoiser: {opt.denoiser}')
    if opt.denoiser == 'unet':
        return UNetDenoiser2D(ckpt_path, state_dict, precision)
    if opt.denoiser == 'ircnn':
        return IRCNNDenoiser2D(ckpt_path, state_dict, precision)
    raise NotImplementedError
