"""The Python autograd nodes of the reverse mode (tfpnp_b200/solver.py: _CSMRIAdmmFn, _SPIAdmmFn, _PRIadmmFn, _CSMRIVariantFn) run
END TO END ON THE CPU: the native forward is replaced by the oracle (solver._run / solver.forward stubs), the native backward by
the g++ emulation of the CUDA sequences (tests/grad_elem_host.cpp) behind a stub of the ctypes library with the real entry points'
argument lists.  What this covers that nothing else can without a GPU: argument order / strides / shapes of the ctypes calls, the
trajectory recording, the widening of the hyper-parameter gradients, and the number and order of the values `backward` returns.
Results are compared with the fixtures recorded from the unmodified reference under autograd.
"""
import contextlib
import ctypes as C
import os
import shutil
import subprocess
import types

import pytest
import torch

from conftest import load_golden, rel_err, weights
from oracle import pnp_oracle as O


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    so = str(tmp_path_factory.mktemp("emu_glue") / "grad_elem_host.so")
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "grad_elem_host.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, src], check=True)
    return C.CDLL(so)


def _vp(x):
    return C.c_void_p(x) if x else None


class _StubLib:
    """The reverse-mode entry points of include/tfpnp_b200.h with their real argument lists, served by the emulation."""

    def __init__(self, emu, flat):
        self.emu, self.flat = emu, flat

    def tfpnp_last_error(self):
        return b"stub"

    def tfpnp_csmri_admm_backward(self, den, states, y0, mask, sg, mu, rs, cs, B, N, iters, gout, gs, gm, gst, stream):
        assert (rs, cs) == (iters, 1) and den == "grad-handle" and stream == 0
        return self.emu.emu_admm_backward(_vp(self.flat.data_ptr()), _vp(states), _vp(y0), _vp(mask), _vp(sg), _vp(mu), B, N, iters,
                                          _vp(gout), _vp(gs), _vp(gm), _vp(gst))

    def tfpnp_spi_admm_backward(self, den, states, x0, K, K_stride, sg, mu, rs, cs, B, H, W, iters, gout, gs, gm, gst, stream):
        assert (rs, cs) == (iters, 1) and den == "grad-handle"
        return self.emu.emu_spi_backward(_vp(self.flat.data_ptr()), _vp(states), _vp(x0), _vp(K), C.c_int64(K_stride), _vp(sg), _vp(mu),
                                         B, H, W, iters, _vp(gout), _vp(gs), _vp(gm), _vp(gst))

    def tfpnp_pr_iadmm_backward(self, den, states, y0, mask, M, sg, mu, tau, rs, cs, B, N, iters, gout, gs, gm, gt, gst, stream):
        assert (rs, cs) == (iters, 1) and den == "grad-handle"
        return self.emu.emu_pr_backward(_vp(self.flat.data_ptr()), _vp(states), _vp(y0), _vp(mask), M, _vp(sg), _vp(mu), _vp(tau), B, N,
                                        iters, _vp(gout), _vp(gs), _vp(gm), _vp(gt), _vp(gst))

    def tfpnp_csmri_variant_backward(self, algo, den, states, y0, mask, p0, p1, p2, rs, cs, B, N, iters, gout, g0, g1, g2, gst, stream):
        assert (rs, cs) == (iters, 1) and den == "grad-handle"
        return self.emu.emu_variant_backward(algo, _vp(self.flat.data_ptr()), _vp(states), _vp(y0), _vp(mask), _vp(p0), _vp(p1), _vp(p2),
                                             B, N, iters, _vp(gout), _vp(g0), _vp(g1), _vp(g2), _vp(gst))


@pytest.fixture
def cpu_native(emu, monkeypatch):
    """Route the autograd nodes' native calls to the CPU stand-ins."""
    from tfpnp_b200 import _lib
    from tfpnp_b200.denoiser import flatten_state_dict
    flat = flatten_state_dict(weights("he"))
    stub = _StubLib(emu, flat)
    monkeypatch.setattr(_lib, "lib", lambda: stub)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a: types.SimpleNamespace(cuda_stream=0))
    return types.SimpleNamespace(_grad_handle=lambda dev: "grad-handle")


def test_csmri_admm_node_on_cpu(cpu_native):
    from tfpnp_b200.solver import _CSMRIAdmmFn
    g = load_golden("grad_csmri_small")
    sd = weights("he")

    def run(handle, state, y0, m8, stride, params, iters):          # _NativeADMM._run
        assert handle == "solver-handle" and stride == 0 and iters == 1
        return O.admm_csmri(sd, state, y0, m8, params[0], params[1])

    solver = types.SimpleNamespace(_run=run, denoiser=cpu_native)
    state = g["state"].clone().requires_grad_(True)
    # wider hyper-parameter tensors than iter_num: the surplus columns must get zero gradient
    pad = torch.full((2, 2), 0.3)
    sg = torch.cat([g["sigma_d"], pad], 1).requires_grad_(True)
    mu = torch.cat([g["mu"], pad], 1).requires_grad_(True)
    out = _CSMRIAdmmFn.apply(solver, "solver-handle", state, g["y0"], g["mask"].to(torch.uint8), sg, mu, 3)
    assert rel_err(out, O.admm_csmri(sd, g["state"], g["y0"], g["mask"], g["sigma_d"], g["mu"]))[1] <= 1e-6
    gs, gm, gst = torch.autograd.grad(out, (sg, mu, state), g["gout"])
    assert gs.shape == sg.shape and torch.count_nonzero(gs[:, 3:]) == 0 and torch.count_nonzero(gm[:, 3:]) == 0
    assert rel_err(gs[:, :3], g["g_sigma_d"])[0] <= 1e-3 and rel_err(gm[:, :3], g["g_mu"])[0] <= 1e-3
    assert rel_err(gst, g["g_state"])[0] <= 1e-3


def test_spi_admm_node_on_cpu(cpu_native):
    from tfpnp_b200.solver import _SPIAdmmFn
    g = load_golden("grad_spi_small")
    sd = weights("he")
    Kv = g["K"][:, 0, 0, 0].contiguous()

    def run(handle, state, x0, K, K_stride, params, iters):
        assert K_stride == 1 and iters == 1
        return O.admm_spi(sd, state, x0, K.reshape(-1, 1, 1, 1).expand(-1, 1, 1, 1), params[0], params[1])

    solver = types.SimpleNamespace(_run=run, denoiser=cpu_native)
    state = g["state"].clone().requires_grad_(True)
    sg = g["sigma_d"].clone().requires_grad_(True)
    mu = g["mu"].clone().requires_grad_(True)
    out = _SPIAdmmFn.apply(solver, "solver-handle", state, g["x0"], Kv, sg, mu, 3)
    gs, gm, gst = torch.autograd.grad(out, (sg, mu, state), g["gout"])
    for mine, key in ((gs, "g_sigma_d"), (gm, "g_mu"), (gst, "g_state")):
        assert rel_err(mine, g[key])[0] <= 1e-3, key


def test_pr_iadmm_node_on_cpu(cpu_native):
    from tfpnp_b200.solver import _PRIadmmFn
    g = load_golden("grad_pr_small")
    sd = weights("he")

    def run(handle, state, y0, mask, stride, params, iters):
        return O.iadmm_pr(sd, state, y0, mask, *params)

    solver = types.SimpleNamespace(_run=run, denoiser=cpu_native)
    state = g["state"].clone().requires_grad_(True)
    ps = [g[k].clone().requires_grad_(True) for k in ("sigma_d", "mu", "tau")]
    out = _PRIadmmFn.apply(solver, "solver-handle", state, g["y0"], g["mask"], *ps, 3)
    grads = torch.autograd.grad(out, (*ps, state), g["gout"])
    for mine, key in zip(grads, ("g_sigma_d", "g_mu", "g_tau", "g_state")):
        assert rel_err(mine, g[key])[0] <= 1e-3, key


@pytest.mark.parametrize("name,algo,keys", [("hqs", 1, ("sigma_d", "mu")), ("apg", 3, ("sigma_d", "tau", "beta"))])
def test_variant_node_on_cpu(cpu_native, name, algo, keys):
    from tfpnp_b200.solver import _CSMRIVariantFn
    g = load_golden("grad_csmri_variants")
    sd = weights("he")
    fn = {"hqs": O.hqs_csmri, "apg": O.apg_csmri}[name]

    def forward(inputs, params, iters):                                # _CSMRIVariant.forward
        state, (y0, m8) = inputs
        assert iters == 1
        return fn(sd, state, y0, m8, *params)

    solver = types.SimpleNamespace(forward=forward, denoiser=cpu_native, _algo=algo)
    state = g[name + "_state0"].clone().requires_grad_(True)
    ps = [g[k].clone().requires_grad_(True) for k in keys]
    out = _CSMRIVariantFn.apply(solver, state, g["y0"], g["mask"].to(torch.uint8), 3, *ps)
    grads = torch.autograd.grad(out, (*ps, state), g[name + "_gout"])
    for mine, k in zip(grads, keys + ("state",)):
        assert rel_err(mine, g[f"{name}_g_{k}"])[0] <= 5e-3, (name, k)


def test_ct_iadmm_node_on_cpu(cpu_native, emu, monkeypatch):
    import numpy as np
    from oracle import grad_oracle as G, synth
    from tfpnp_b200 import _lib
    from tfpnp_b200.denoiser import flatten_state_dict
    from tfpnp_b200.solver import _CTIadmmFn
    sd = weights("he")
    B, N, views, it = 2, 32, 12, 2
    d = synth.ct_batch(B, N, views, it)
    opnorm = float(d["opnorm"])
    cs, sn, det = O.ct_geometry(N, views)
    flat = flatten_state_dict(sd)

    @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)
    def ata(img_p, with_y0, out_p):
        img = torch.from_numpy(np.ctypeslib.as_array(C.cast(img_p, C.POINTER(C.c_float)), shape=(B, 1, N, N)).copy())
        s = O.radon_forward(img, cs, sn, det)
        if with_y0:
            s = s - d["y0"]
        w = O.radon_backward(s, cs, sn, N).contiguous().float()
        C.memmove(out_p, w.data_ptr(), w.numel() * 4)
        return 0

    class Stub:
        def tfpnp_last_error(self):
            return b"stub"

        def tfpnp_ct_iadmm_backward(self, den, states, y0, n_views, op, cos, sin, sg, mu, tau, rs, cs_, Bn, Nn, iters, gout, gs, gm, gt, gst, stream):
            assert (rs, cs_) == (iters, 1) and den == "grad-handle" and n_views == views and abs(op - opnorm) < 1e-6
            assert cos == "cos-ptr" and sin == "sin-ptr"
            return emu.emu_ct_backward(_vp(flat.data_ptr()), _vp(states), ata, C.c_float(op), _vp(sg), _vp(mu), _vp(tau), Bn, Nn, iters,
                                       _vp(gout), _vp(gs), _vp(gm), _vp(gt), _vp(gst))

    monkeypatch.setattr(_lib, "lib", lambda: Stub())

    def run(handle, state, y0, aux1, stride, params, iters):
        assert aux1 is None and iters == 1
        return O.iadmm_ct(sd, state, y0, views, opnorm, *params)

    solver = types.SimpleNamespace(_run=run, denoiser=cpu_native)
    g = torch.Generator().manual_seed(23)
    state0 = torch.cat([torch.rand(B, 1, N, N, generator=g), torch.rand(B, 1, N, N, generator=g),
                        torch.rand(B, 1, N, N, generator=g) * 0.3], dim=1)
    gout = torch.randn(state0.shape, generator=g)
    ref = G.iadmm_ct_vjp_autograd(sd, state0, d["y0"], views, opnorm, d["sigma_d"], d["mu"], d["tau"], gout)
    state = state0.clone().requires_grad_(True)
    ps = [d[k].clone().requires_grad_(True) for k in ("sigma_d", "mu", "tau")]
    fake_table = types.SimpleNamespace(data_ptr=lambda: "cos-ptr")
    fake_table2 = types.SimpleNamespace(data_ptr=lambda: "sin-ptr")
    out = _CTIadmmFn.apply(solver, "solver-handle", state, d["y0"], *ps, it, views, opnorm, fake_table, fake_table2)
    grads = torch.autograd.grad(out, (*ps, state), gout)
    for mine, r, name in zip(grads, ref, ("sigma_d", "mu", "tau", "state")):
        assert rel_err(mine, r)[0] <= 1e-3, (name, rel_err(mine, r))
