"""B200-native PnP solvers behind the reference's ``PnPSolver`` interface.

Mirrors tfpnp/pnp/solver/base.py:5-116 and the four task solvers
(tasks/csmri/solver.py:9-57, tasks/pr/solver.py:15-76, tasks/ct/solver.py:7-53,
tasks/spi/solver.py:8-51): same class names, ``reset / forward(inputs, parameters,
iter_num=None) / get_output / prox_mapping / num_var / filter_aux_inputs /
filter_hyperparameter``, so ``PnPEnv.step`` (tfpnp/env/base.py:157-191) drives them
unchanged.  ``forward`` marshals raw device pointers into ``tfpnp_solver_forward``
(include/tfpnp_b200.h); the iterated proximal loop itself is hand-written sm_100a CUDA.

There is no PyTorch/CPU fallback.  The differentiable use (``PnPEnv.forward`` under
autograd, tfpnp/env/base.py:193-206) exists for the four ADMM / iADMM solvers
(gradients w.r.t. the hyper-parameters and the input state through
``tfpnp_{csmri_admm,pr_iadmm,ct_iadmm,spi_admm}_backward``) and for the HQS / PG / APG / RED-ADMM
CS-MRI solvers (``tfpnp_csmri_variant_backward``); it is on by default (``solver.differentiable``)
and validated on a B200 (tests/test_grad.py); PGSolver_CT raises NotImplementedError under autograd.
"""
from __future__ import annotations

import ctypes as C
import math

import torch
import torch.nn as nn

from . import _lib
from .denoiser import UNetDenoiser2D


class PnPSolver(nn.Module):
    """tfpnp/pnp/solver/base.py:5-84."""

    def __init__(self, denoiser):
        super().__init__()
        self.denoiser = denoiser

    def reset(self, data):
        raise NotImplementedError

    def forward(self, inputs, parameters, iter_num):
        raise NotImplementedError

    def get_output(self, state):
        raise NotImplementedError

    def prox_mapping(self, x, sigma):
        return self.denoiser(x, sigma)

    @property
    def num_var(self):
        raise NotImplementedError

    def filter_aux_inputs(self, state):
        raise NotImplementedError

    def filter_hyperparameter(self, action):
        raise NotImplementedError


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()


def _expect(name, t, shape):
    """Raise ValueError unless tensor ``t`` has exactly ``shape`` (the reference fails in its first broadcast; raw
    pointers would read out of bounds instead)."""
    if tuple(t.shape) != tuple(shape):
        raise ValueError(f"{name} must have shape {tuple(shape)}, got {tuple(t.shape)}")


def _expect_params(params, B):
    for i, p in enumerate(params):
        if p.dim() < 1 or p.shape[0] != B:
            raise ValueError(f"hyper-parameter {i} must have a leading dimension of {B}, got {tuple(p.shape)}")
        if not p.is_cuda:
            raise RuntimeError("tfpnp_b200 solvers run on CUDA (sm_100) tensors only; there is no CPU fallback")


class _NativeADMM(PnPSolver):
    """Shared host logic: handle cache + pointer marshalling for one task."""
    _task = None
    _complex_state = False
    use_graph = True

    def __init__(self, denoiser):
        if not isinstance(denoiser, UNetDenoiser2D):      # IRCNNDenoiser2D derives from it
            raise TypeError("tfpnp_b200 solvers need a tfpnp_b200.UNetDenoiser2D / IRCNNDenoiser2D (the denoiser "
                            "runs inside the fused CUDA path)")
        super().__init__(denoiser)
        self._solvers = {}      # (device idx, H, W, extra, denoiser engine) -> handle
        self._den_refs = []     # denoisers whose native engine a cached handle captured
        self.last_launch_count = 0

    @property
    def num_var(self):                      # base.py:91-93
        return 3

    def reset(self, data):                  # base.py:95-99
        x = data['x0'].clone().detach()
        z = x.clone().detach()
        u = torch.zeros_like(x)
        return torch.cat((x, z, u), dim=1)

    def get_output(self, state):            # base.py:101-104
        x, _, _ = torch.split(state, state.shape[1] // 3, dim=1)
        return x[..., 0] if self._complex_state else x

    def filter_hyperparameter(self, action):  # base.py:106-107
        return action['sigma_d'], action['mu']

    # -- native plumbing -------------------------------------------------------
    def _solver_handle(self, device, H, W, n_masks=0, views=0, opnorm=0.0, cos=None, sin=None):
        idx = device.index if device.index is not None else torch.cuda.current_device()
        den_h = self.denoiser._handle(device)
        # keyed on the native denoiser engine too: re-assigning solver.denoiser must not leave a cached solver
        # handle pointing at the old (possibly destroyed) engine; _den_refs keeps every engine we captured alive
        key = (idx, H, W, n_masks, views, float(opnorm), den_h.value)
        h = self._solvers.get(key)
        if h is None:
            self._den_refs.append(self.denoiser)
            cfg = _lib.SolverConfig(self._task, H, W, n_masks, views, float(opnorm), None, None,
                                    1 if self.use_graph else 0)
            if cos is not None:
                cfg.ct_cos = C.cast(cos.data_ptr(), C.POINTER(C.c_float))
                cfg.ct_sin = C.cast(sin.data_ptr(), C.POINTER(C.c_float))
            out = C.c_void_p()
            with torch.cuda.device(idx):
                _lib.check(_lib.lib().tfpnp_solver_create(C.byref(cfg), den_h, C.byref(out)), "tfpnp_solver_create")
            self._solvers[key] = h = out
        return h

    # reverse mode: the four ADMM / iADMM solvers and the CS-MRI variants define a native backward;
    # ``differentiable = False`` refuses gradient requests loudly
    differentiable = True
    _has_backward = False

    def _wants_grad(self, variables, parameters):
        return torch.is_grad_enabled() and (variables.requires_grad or any(p.requires_grad for p in parameters))

    def _check_inputs(self, variables, parameters):
        if not variables.is_cuda:
            raise RuntimeError("tfpnp_b200 solvers run on CUDA (sm_100) tensors only; there is no CPU fallback")
        if self._wants_grad(variables, parameters) and not (self.differentiable and self._has_backward):
            raise NotImplementedError(
                "the differentiable solver path (PnPEnv.forward under autograd, SURVEY 8f N4) is built for the four "
                "ADMM / iADMM solvers and the CS-MRI variants, and this solver has it switched off or lacks it")

    def _run(self, handle, variables, aux0, aux1, aux1_stride, params, iter_num):
        B = variables.shape[0]
        sigma_d = params[0]
        if iter_num is None:                # infer from the hyper-parameters (tasks/csmri/solver.py:40-41)
            iter_num = sigma_d.shape[-1]
        state_in = _f32c(variables)
        out = torch.empty_like(state_in)
        # all parameter tensors must share strides; slices of one action tensor do, else copy
        ps = [p if p.dtype == torch.float32 else p.float() for p in params]
        ps = [p.reshape(B, -1) for p in ps]
        if any(p.stride() != ps[0].stride() for p in ps):
            ps = [p.contiguous() for p in ps]
        if any(p.shape[1] < iter_num for p in ps):
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width {ps[0].shape[1]}")
        rs, cs = ps[0].stride()
        tau_ptr = ps[2].data_ptr() if len(ps) > 2 else None
        with torch.cuda.device(variables.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(_lib.lib().tfpnp_solver_forward(
                handle, state_in.data_ptr(), aux0.data_ptr(), aux1.data_ptr() if aux1 is not None else None,
                aux1_stride, ps[0].data_ptr(), ps[1].data_ptr(), tau_ptr, rs, cs, B, int(iter_num),
                out.data_ptr(), st), "tfpnp_solver_forward")
            self.last_launch_count = int(_lib.lib().tfpnp_solver_last_launch_count(handle))
        return out

    def __del__(self):
        try:                                # solver handles first: they borrow the denoiser engines in _den_refs
            for h in self._solvers.values():
                _lib.lib().tfpnp_solver_destroy(h)
            self._solvers.clear()
        except Exception:
            pass


class ADMMSolver(_NativeADMM):
    """tfpnp/pnp/solver/base.py:87-107."""


class IADMMSolver(ADMMSolver):
    """tfpnp/pnp/solver/base.py:110-116 (inexact ADMM: adds tau)."""

    def filter_hyperparameter(self, action):
        return action['sigma_d'], action['mu'], action['tau']


class ADMMSolver_CSMRI(ADMMSolver):
    """tasks/csmri/solver.py:9-57."""
    _task = _lib.TASK_CSMRI
    _complex_state = True
    _has_backward = True

    def filter_aux_inputs(self, state):     # solver.py:20-21
        return (state['y0'], state['mask'])

    def forward(self, inputs, parameters, iter_num=None):
        variables, aux = inputs
        y0, mask = tuple(aux)               # may be a one-shot generator (tfpnp/utils/misc.py:138-139)
        sigma_d, mu = parameters
        self._check_inputs(variables, (sigma_d, mu))
        if variables.dim() != 5:
            raise ValueError(f"variables must be [B,3,H,W,2], got {tuple(variables.shape)}")
        B, _, H, W, _ = variables.shape
        _expect("variables", variables, (B, 3, H, W, 2))
        _expect("y0", y0, (B, 1, H, W, 2))
        _expect("mask", mask, (B, 1, H, W))
        _expect_params((sigma_d, mu), B)
        if H != W:
            raise ValueError(f"square images only, got {H}x{W}")
        m8 = mask.contiguous()
        m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else (m8 != 0).view(torch.uint8)
        h = self._solver_handle(variables.device, H, W)
        if self._wants_grad(variables, (sigma_d, mu)):
            if iter_num is None:
                iter_num = sigma_d.shape[-1]
            return _CSMRIAdmmFn.apply(self, h, variables, _f32c(y0), m8, sigma_d, mu, int(iter_num))
        return self._run(h, variables, _f32c(y0), m8, 0, (sigma_d, mu), iter_num)


class _CSMRIAdmmFn(torch.autograd.Function):
    """autograd node for ADMMSolver_CSMRI.forward (what tfpnp/trainer/mddpg/trainer.py:173 differentiates).

    forward: the native solver one iteration at a time, recording the trajectory; backward: the adjoint recursion
    of tfpnp_csmri_admm_backward (csmri_variants.cu) with the denoiser's vector-Jacobian product on the fp32 engine.
    """

    @staticmethod
    def forward(ctx, solver, handle, variables, y0, m8, sigma_d, mu, iter_num):
        B = variables.shape[0]
        sg = sigma_d.detach().float().reshape(B, -1)[:, :iter_num].contiguous()
        m = mu.detach().float().reshape(B, -1)[:, :iter_num].contiguous()
        if sg.shape[1] < iter_num or m.shape[1] < iter_num:
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width")
        states = [_f32c(variables.detach())]
        with torch.no_grad():
            for i in range(iter_num):
                states.append(solver._run(handle, states[-1], y0, m8, 0, (sg[:, i:i + 1], m[:, i:i + 1]), 1))
        ctx.solver = solver
        ctx.meta = (sigma_d.shape, mu.shape, sigma_d.dtype, mu.dtype, iter_num)
        ctx.save_for_backward(torch.stack(states), y0, m8, sg, m)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, gout):
        states, y0, m8, sg, m = ctx.saved_tensors
        sshape, mshape, sdt, mdt, it = ctx.meta
        B, _, H, W, _ = gout.shape
        gout = _f32c(gout)
        g_sigma = torch.zeros(B, it, device=gout.device, dtype=torch.float32)
        g_mu = torch.zeros_like(g_sigma)
        g_state = torch.empty_like(gout)
        den = ctx.solver.denoiser
        with torch.cuda.device(gout.device):
            _lib.check(_lib.lib().tfpnp_csmri_admm_backward(
                den._grad_handle(gout.device), states.data_ptr(), y0.data_ptr(), m8.data_ptr(), sg.data_ptr(),
                m.data_ptr(), it, 1, B, H, it, gout.data_ptr(), g_sigma.data_ptr(), g_mu.data_ptr(), g_state.data_ptr(),
                torch.cuda.current_stream().cuda_stream), "tfpnp_csmri_admm_backward")

        def widen(g, shape, dtype):          # parameters beyond iter_num did not take part: zero gradient
            full = torch.zeros(B, max(1, math.prod(shape[1:])), device=g.device, dtype=torch.float32)
            full[:, :it] = g
            return full.reshape(shape).to(dtype)

        return (None, None, g_state, None, None, widen(g_sigma, sshape, sdt), widen(g_mu, mshape, mdt), None)


class _CSMRIVariant(PnPSolver):
    """Host logic shared by the other CS-MRI solvers (tasks/csmri/solver.py:60-204): same kernels, different update."""
    _algo = None
    _nvar = None
    _param_keys = ()
    differentiable = True       # reverse mode; False refuses gradient requests loudly

    def __init__(self, denoiser):
        if not isinstance(denoiser, UNetDenoiser2D):
            raise TypeError("tfpnp_b200 solvers need a tfpnp_b200.UNetDenoiser2D / IRCNNDenoiser2D")
        super().__init__(denoiser)
        self._solvers = {}
        self._den_refs = []
        self.last_launch_count = 0

    @property
    def num_var(self):
        return self._nvar

    def get_output(self, state):            # CSMRIMixin.get_output: complex2real of the first variable
        x = torch.split(state, state.shape[1] // self._nvar, dim=1)[0]
        return x[..., 0]

    def filter_aux_inputs(self, state):     # solver.py:20-21
        return (state['y0'], state['mask'])

    def filter_hyperparameter(self, action):
        return tuple(action[k] for k in self._param_keys)

    def forward(self, inputs, parameters, iter_num=None):
        variables, aux = inputs
        y0, mask = tuple(aux)
        params = tuple(parameters)
        if not variables.is_cuda:
            raise RuntimeError("tfpnp_b200 solvers run on CUDA (sm_100) tensors only; there is no CPU fallback")
        wants_grad = torch.is_grad_enabled() and (variables.requires_grad or any(p.requires_grad for p in params))
        if wants_grad and not self.differentiable:
            raise NotImplementedError("the differentiable solver path was switched off (solver.differentiable = False)")
        if variables.dim() != 5:
            raise ValueError(f"variables must be [B,{self._nvar},H,W,2], got {tuple(variables.shape)}")
        B, _, H, W, _ = variables.shape
        _expect("variables", variables, (B, self._nvar, H, W, 2))
        _expect("y0", y0, (B, 1, H, W, 2))
        _expect("mask", mask, (B, 1, H, W))
        _expect_params(params, B)
        if H != W:
            raise ValueError(f"square images only, got {H}x{W}")
        if len(params) != len(self._param_keys):
            raise ValueError(f"{type(self).__name__} takes {len(self._param_keys)} hyper-parameters {self._param_keys}")
        if iter_num is None:
            iter_num = params[0].shape[-1]
        if wants_grad:
            m8 = mask.contiguous()
            m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else (m8 != 0).view(torch.uint8)
            return _CSMRIVariantFn.apply(self, variables, _f32c(y0), m8, int(iter_num), *params)
        dev = variables.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        den_h = self.denoiser._handle(dev)
        h = self._solvers.get((idx, H, den_h.value))
        if h is None:
            h = C.c_void_p()
            self._den_refs.append(self.denoiser)
            with torch.cuda.device(idx):
                _lib.check(_lib.lib().tfpnp_csmri_variant_create(self._algo, H, den_h, C.byref(h)),
                           "tfpnp_csmri_variant_create")
            self._solvers[(idx, H, den_h.value)] = h
        m8 = mask.contiguous()
        m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else (m8 != 0).view(torch.uint8)
        ps = [(p if p.dtype == torch.float32 else p.float()).reshape(B, -1) for p in params]
        if any(p.stride() != ps[0].stride() for p in ps):
            ps = [p.contiguous() for p in ps]
        if any(p.shape[1] < iter_num for p in ps):
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width {ps[0].shape[1]}")
        rs, cs = ps[0].stride()
        state_in = _f32c(variables)
        out = torch.empty_like(state_in)
        y0c = _f32c(y0)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib().tfpnp_csmri_variant_forward(
                h, state_in.data_ptr(), y0c.data_ptr(), m8.data_ptr(), ps[0].data_ptr(), ps[1].data_ptr(),
                ps[2].data_ptr() if len(ps) > 2 else None, rs, cs, B, int(iter_num), out.data_ptr(),
                torch.cuda.current_stream().cuda_stream), "tfpnp_csmri_variant_forward")
        return out

    def __del__(self):
        try:
            for h in self._solvers.values():
                _lib.lib().tfpnp_csmri_variant_destroy(h)
        except Exception:
            pass


class _CSMRIVariantFn(torch.autograd.Function):
    """autograd node for the HQS / PG / APG / RED-ADMM solvers: trajectory with the native forward (one iteration per call),
    tfpnp_csmri_variant_backward (csmri_variants.cu)."""

    @staticmethod
    def forward(ctx, solver, variables, y0, m8, iter_num, *params):
        B = variables.shape[0]
        ps = [p.detach().float().reshape(B, -1)[:, :iter_num].contiguous() for p in params]
        if any(p.shape[1] < iter_num for p in ps):
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width")
        states = [_f32c(variables.detach())]
        with torch.no_grad():
            for i in range(iter_num):
                states.append(solver.forward((states[-1], (y0, m8)), tuple(p[:, i:i + 1] for p in ps), 1))
        ctx.solver = solver
        ctx.meta = ([p.shape for p in params], [p.dtype for p in params], iter_num)
        ctx.save_for_backward(torch.stack(states), y0, m8, *ps)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, gout):
        states, y0, m8, *ps = ctx.saved_tensors
        shapes, dtypes, it = ctx.meta
        B, _, H, W, _ = gout.shape
        gout = _f32c(gout)
        grads = [torch.zeros(B, it, device=gout.device, dtype=torch.float32) for _ in ps]
        g_state = torch.empty_like(gout)
        ptr = lambda k, arr: arr[k].data_ptr() if k < len(arr) else None
        with torch.cuda.device(gout.device):
            _lib.check(_lib.lib().tfpnp_csmri_variant_backward(
                ctx.solver._algo, ctx.solver.denoiser._grad_handle(gout.device), states.data_ptr(), y0.data_ptr(), m8.data_ptr(),
                ptr(0, ps), ptr(1, ps), ptr(2, ps), it, 1, B, H, it, gout.data_ptr(), ptr(0, grads), ptr(1, grads), ptr(2, grads),
                g_state.data_ptr(), torch.cuda.current_stream().cuda_stream), "tfpnp_csmri_variant_backward")

        def widen(g, shape, dtype):
            full = torch.zeros(B, max(1, math.prod(shape[1:])), device=g.device, dtype=torch.float32)
            full[:, :it] = g
            return full.reshape(shape).to(dtype)

        return (None, g_state, None, None, None, *[widen(g, sh, dt) for g, sh, dt in zip(grads, shapes, dtypes)])


class HQSSolver_CSMRI(_CSMRIVariant):
    """tasks/csmri/solver.py:60-88 + HQSSolver (tfpnp/pnp/solver/base.py:118-139)."""
    _algo, _nvar, _param_keys = _lib.ALGO_HQS, 2, ('sigma_d', 'mu')

    def reset(self, data):
        x = data['x0'].clone().detach()
        return torch.cat([x, x.clone().detach()], dim=1)


class PGSolver_CSMRI(_CSMRIVariant):
    """tasks/csmri/solver.py:91-118 + PGSolver (base.py:141-160)."""
    _algo, _nvar, _param_keys = _lib.ALGO_PG, 1, ('sigma_d', 'tau')

    def reset(self, data):
        return data['x0'].clone().detach()


class APGSolver_CSMRI(_CSMRIVariant):
    """tasks/csmri/solver.py:121-161 + APGSolver (base.py:162-189); beta comes from the action, as upstream."""
    _algo, _nvar, _param_keys = _lib.ALGO_APG, 2, ('sigma_d', 'tau', 'beta')

    def reset(self, data):
        x = data['x0'].clone().detach()
        return torch.cat([x, x.clone().detach()], dim=1)


class REDADMMSolver_CSMRI(_CSMRIVariant):
    """tasks/csmri/solver.py:164-201 + REDADMMSolver (base.py:192-214)."""
    _algo, _nvar, _param_keys = _lib.ALGO_REDADMM, 3, ('sigma_d', 'mu', 'lamda')

    def reset(self, data):
        x = data['x0'].clone().detach()
        return torch.cat([x, x.clone().detach(), torch.zeros_like(x)], dim=1)


class IADMMSolver_PR(IADMMSolver):
    """tasks/pr/solver.py:15-76."""
    _task = _lib.TASK_PR
    _complex_state = True
    _has_backward = True

    def filter_aux_inputs(self, state):     # solver.py:20-21
        return (state['y0'], state['mask'])

    def reset(self, data):                  # solver.py:29-35
        x0 = data['x0'].clone().detach()
        x = torch.stack([x0, torch.zeros_like(x0)], dim=4)
        return torch.cat([x, x.clone(), torch.zeros_like(x)], dim=1)

    def forward(self, inputs, parameters, iter_num=None):
        variables, aux = inputs
        y0, mask = tuple(aux)
        sigma_d, mu, tau = parameters
        self._check_inputs(variables, (sigma_d, mu, tau))
        if variables.dim() != 5 or mask.dim() != 5:
            raise ValueError(f"variables must be [B,3,H,W,2] and mask [B,M,H,W,2], got {tuple(variables.shape)}, {tuple(mask.shape)}")
        B, _, H, W, _ = variables.shape
        M = mask.shape[1]
        _expect("variables", variables, (B, 3, H, W, 2))
        _expect("mask", mask, (B, M, H, W, 2))
        _expect("y0", y0, (B, M, H, W))
        _expect_params((sigma_d, mu, tau), B)
        if H != W:
            raise ValueError(f"square images only, got {H}x{W}")
        h = self._solver_handle(variables.device, H, W, n_masks=M)
        if self._wants_grad(variables, (sigma_d, mu, tau)):
            if iter_num is None:
                iter_num = sigma_d.shape[-1]
            return _PRIadmmFn.apply(self, h, variables, _f32c(y0), _f32c(mask), sigma_d, mu, tau, int(iter_num))
        return self._run(h, variables, _f32c(y0), _f32c(mask), 0, (sigma_d, mu, tau), iter_num)


class _PRIadmmFn(torch.autograd.Function):
    """autograd node for IADMMSolver_PR.forward: trajectory with the native forward, tfpnp_pr_iadmm_backward (pr.cu)."""

    @staticmethod
    def forward(ctx, solver, handle, variables, y0, mask, sigma_d, mu, tau, iter_num):
        B = variables.shape[0]
        ps = [p.detach().float().reshape(B, -1)[:, :iter_num].contiguous() for p in (sigma_d, mu, tau)]
        if any(p.shape[1] < iter_num for p in ps):
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width")
        states = [_f32c(variables.detach())]
        with torch.no_grad():
            for i in range(iter_num):
                states.append(solver._run(handle, states[-1], y0, mask, 0, tuple(p[:, i:i + 1] for p in ps), 1))
        ctx.solver = solver
        ctx.meta = ([p.shape for p in (sigma_d, mu, tau)], [p.dtype for p in (sigma_d, mu, tau)], iter_num)
        ctx.save_for_backward(torch.stack(states), y0, mask, *ps)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, gout):
        states, y0, mask, sg, m, t = ctx.saved_tensors
        shapes, dtypes, it = ctx.meta
        B, _, H, W, _ = gout.shape
        gout = _f32c(gout)
        grads = [torch.zeros(B, it, device=gout.device, dtype=torch.float32) for _ in range(3)]
        g_state = torch.empty_like(gout)
        with torch.cuda.device(gout.device):
            _lib.check(_lib.lib().tfpnp_pr_iadmm_backward(
                ctx.solver.denoiser._grad_handle(gout.device), states.data_ptr(), y0.data_ptr(), mask.data_ptr(), mask.shape[1],
                sg.data_ptr(), m.data_ptr(), t.data_ptr(), it, 1, B, W, it, gout.data_ptr(), grads[0].data_ptr(),
                grads[1].data_ptr(), grads[2].data_ptr(), g_state.data_ptr(), torch.cuda.current_stream().cuda_stream),
                "tfpnp_pr_iadmm_backward")

        def widen(g, shape, dtype):
            full = torch.zeros(B, max(1, math.prod(shape[1:])), device=g.device, dtype=torch.float32)
            full[:, :it] = g
            return full.reshape(shape).to(dtype)

        gs, gm, gt = (widen(g, sh, dt) for g, sh, dt in zip(grads, shapes, dtypes))
        return (None, None, g_state, None, None, gs, gm, gt, None)


class RadonGenerator:
    """Per-(resolution, views) operator-norm cache, as tfpnp/utils/transforms.py:494-508.
    The power method (transforms.py:447-462) runs on the GPU operators of libtfpnp_b200 from a
    SEEDED start vector (the reference starts from an unseeded torch.randn)."""

    def __init__(self):
        self.opnorms = {}

    @staticmethod
    def tables(views):
        angles = torch.linspace(0, 179 / 180 * math.pi, views, dtype=torch.float32)   # transforms.py:488
        return torch.cos(angles.double()).float().contiguous(), torch.sin(angles.double()).float().contiguous()

    def __call__(self, resolution, views, device):
        key = (resolution, views)
        if key not in self.opnorms:
            from .ops import radon_forward, radon_backward
            g = torch.Generator().manual_seed(0)
            x = torch.randn(1, 1, resolution, resolution, generator=g).to(device)
            x = x / x.norm()
            v = 0.0
            for _ in range(10):
                x = radon_backward(radon_forward(x, views), resolution, views)
                v = float(x.norm())
                x = x / v
            self.opnorms[key] = v ** 0.5
        return self.opnorms[key]


class IADMMSolver_CT(IADMMSolver):
    """tasks/ct/solver.py:7-53 (Radon pair: this build's own discretisation, parity unpinned
    w.r.t. the absent torch_radon)."""
    _task = _lib.TASK_CT
    _has_backward = True

    def __init__(self, denoiser):
        super().__init__(denoiser)
        self.radon_generator = RadonGenerator()
        self.opnorm_override = None         # tests pass the CPU checker's opnorm explicitly

    def filter_aux_inputs(self, state):     # solver.py:8-9
        return (state['y0'], state['view'])

    def forward(self, inputs, parameters, iter_num=None):
        variables, aux = inputs
        y0, view = tuple(aux)
        sigma_d, mu, tau = parameters
        self._check_inputs(variables, (sigma_d, mu, tau))
        if variables.dim() != 4:
            raise ValueError(f"variables must be [B,3,H,W], got {tuple(variables.shape)}")
        B, _, H, W = variables.shape
        _expect("variables", variables, (B, 3, H, W))
        _expect_params((sigma_d, mu, tau), B)
        if H != W:
            raise ValueError(f"square images only, got {H}x{W}")
        views = int(view[0, 0, 0, 0].item() * 120)          # solver.py:26 (one host sync per call)
        _expect("y0", y0, (B, 1, views, math.ceil(math.sqrt(2.0) * W)))   # det_count, transforms.py:489
        opnorm = self.opnorm_override or self.radon_generator(W, views, variables.device)
        cos, sin = RadonGenerator.tables(views)
        h = self._solver_handle(variables.device, H, W, views=views, opnorm=opnorm, cos=cos, sin=sin)
        if self._wants_grad(variables, (sigma_d, mu, tau)):
            if iter_num is None:
                iter_num = sigma_d.shape[-1]
            return _CTIadmmFn.apply(self, h, variables, _f32c(y0), sigma_d, mu, tau, int(iter_num), views, float(opnorm), cos, sin)
        return self._run(h, variables, _f32c(y0), None, 0, (sigma_d, mu, tau), iter_num)


class _CTIadmmFn(torch.autograd.Function):
    """autograd node for IADMMSolver_CT.forward: trajectory with the native forward, tfpnp_ct_iadmm_backward (misc.cu)."""

    @staticmethod
    def forward(ctx, solver, handle, variables, y0, sigma_d, mu, tau, iter_num, views, opnorm, cos, sin):
        B = variables.shape[0]
        ps = [p.detach().float().reshape(B, -1)[:, :iter_num].contiguous() for p in (sigma_d, mu, tau)]
        if any(p.shape[1] < iter_num for p in ps):
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width")
        states = [_f32c(variables.detach())]
        with torch.no_grad():
            for i in range(iter_num):
                states.append(solver._run(handle, states[-1], y0, None, 0, tuple(p[:, i:i + 1] for p in ps), 1))
        ctx.solver = solver
        ctx.meta = ([p.shape for p in (sigma_d, mu, tau)], [p.dtype for p in (sigma_d, mu, tau)], iter_num, views, opnorm, cos, sin)
        ctx.save_for_backward(torch.stack(states), y0, *ps)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, gout):
        states, y0, sg, m, t = ctx.saved_tensors
        shapes, dtypes, it, views, opnorm, cos, sin = ctx.meta
        B, _, H, W = gout.shape
        gout = _f32c(gout)
        grads = [torch.zeros(B, it, device=gout.device, dtype=torch.float32) for _ in range(3)]
        g_state = torch.empty_like(gout)
        with torch.cuda.device(gout.device):
            _lib.check(_lib.lib().tfpnp_ct_iadmm_backward(
                ctx.solver.denoiser._grad_handle(gout.device), states.data_ptr(), y0.data_ptr(), views, opnorm,
                cos.data_ptr(), sin.data_ptr(), sg.data_ptr(), m.data_ptr(), t.data_ptr(), it, 1, B, W, it, gout.data_ptr(),
                grads[0].data_ptr(), grads[1].data_ptr(), grads[2].data_ptr(), g_state.data_ptr(),
                torch.cuda.current_stream().cuda_stream), "tfpnp_ct_iadmm_backward")

        def widen(g, shape, dtype):
            full = torch.zeros(B, max(1, math.prod(shape[1:])), device=g.device, dtype=torch.float32)
            full[:, :it] = g
            return full.reshape(shape).to(dtype)

        gs, gm, gt = (widen(g, sh, dt) for g, sh, dt in zip(grads, shapes, dtypes))
        return (None, None, g_state, None, gs, gm, gt, None, None, None, None, None)


class PGSolver_CT(PnPSolver):
    """tasks/ct/solver.py:56-87 + PGSolver (tfpnp/pnp/solver/base.py:141-160): proximal gradient with this build's
    Radon pair.  A composition of the native operators (projector, backprojector, denoiser); the AXPYs around them are
    torch element-wise glue.  (PGSolver_PR, tasks/pr/solver.py:79-112, applies the CS-MRI gradient with a boolean
    index on the complex CDP masks and cannot run upstream: not built.)"""

    def __init__(self, denoiser):
        if not isinstance(denoiser, UNetDenoiser2D):
            raise TypeError("tfpnp_b200 solvers need a tfpnp_b200.UNetDenoiser2D / IRCNNDenoiser2D")
        super().__init__(denoiser)
        self.radon_generator = RadonGenerator()
        self.opnorm_override = None

    @property
    def num_var(self):
        return 1

    def reset(self, data):
        return data['x0'].clone().detach()

    def get_output(self, state):
        return state

    def filter_aux_inputs(self, state):
        return (state['y0'], state['view'])

    def filter_hyperparameter(self, action):
        return action['sigma_d'], action['tau']

    def forward(self, inputs, parameters, iter_num=None):
        from .ops import radon_forward, radon_backward
        variables, aux = inputs
        y0, view = tuple(aux)
        sigma_d, tau = parameters
        if not variables.is_cuda:
            raise RuntimeError("tfpnp_b200 solvers run on CUDA (sm_100) tensors only; there is no CPU fallback")
        if torch.is_grad_enabled() and (variables.requires_grad or sigma_d.requires_grad or tau.requires_grad):
            raise NotImplementedError("the differentiable solver path is out of scope (SURVEY 8f N4)")
        x = variables
        B, n = x.shape[0], x.shape[-1]
        views = int(view[0, 0, 0, 0].item() * 120)          # solver.py:67
        opnorm = self.opnorm_override or self.radon_generator(n, views, x.device)
        if iter_num is None:
            iter_num = sigma_d.shape[-1]
        for i in range(iter_num):
            _tau = tau[:, i].reshape(B, 1, 1, 1)
            z = x - _tau * (radon_backward(radon_forward(x, views) - y0, n, views) / opnorm ** 2)   # solver.py:79
            x = self.prox_mapping(z, sigma_d[:, i])                                                 # solver.py:82
        return x


class ADMMSolver_SPI(ADMMSolver):
    """tasks/spi/solver.py:8-51."""
    _task = _lib.TASK_SPI
    _has_backward = True

    def filter_aux_inputs(self, state):     # solver.py:9-10
        return (state['x0'], state['K'])

    def forward(self, inputs, parameters, iter_num=None):
        variables, aux = inputs
        x0, K = tuple(aux)
        sigma_d, mu = parameters
        self._check_inputs(variables, (sigma_d, mu))
        if variables.dim() != 4:
            raise ValueError(f"variables must be [B,3,H,W], got {tuple(variables.shape)}")
        B, _, H, W = variables.shape
        _expect("variables", variables, (B, 3, H, W))
        _expect("x0", x0, (B, 1, H, W))
        _expect_params((sigma_d, mu), B)
        if K.dim() != 4 or K.shape[0] != B:
            raise ValueError(f"K must be [B,1,H,W] (constant K/10 per image), got {tuple(K.shape)}")
        Kf = K if K.dtype == torch.float32 else K.float()
        Kv = Kf[:, 0, 0, 0]                                   # solver.py:32 (the *10 happens on device)
        h = self._solver_handle(variables.device, H, W)
        if self._wants_grad(variables, (sigma_d, mu)):
            if iter_num is None:
                iter_num = sigma_d.shape[-1]
            return _SPIAdmmFn.apply(self, h, variables, _f32c(x0), Kv.contiguous(), sigma_d, mu, int(iter_num))
        return self._run(h, variables, _f32c(x0), Kv, Kv.stride(0), (sigma_d, mu), iter_num)


class _SPIAdmmFn(torch.autograd.Function):
    """autograd node for ADMMSolver_SPI.forward: trajectory with the native forward, tfpnp_spi_admm_backward (spi.cu).
    Only the closed-form branch of spi_inverse carries a gradient, as under autograd in the reference."""

    @staticmethod
    def forward(ctx, solver, handle, variables, x0, Kv, sigma_d, mu, iter_num):
        B = variables.shape[0]
        sg = sigma_d.detach().float().reshape(B, -1)[:, :iter_num].contiguous()
        m = mu.detach().float().reshape(B, -1)[:, :iter_num].contiguous()
        if sg.shape[1] < iter_num or m.shape[1] < iter_num:
            raise IndexError(f"iter_num={iter_num} exceeds the hyper-parameter width")
        states = [_f32c(variables.detach())]
        with torch.no_grad():
            for i in range(iter_num):
                states.append(solver._run(handle, states[-1], x0, Kv, 1, (sg[:, i:i + 1], m[:, i:i + 1]), 1))
        ctx.solver = solver
        ctx.meta = (sigma_d.shape, mu.shape, sigma_d.dtype, mu.dtype, iter_num)
        ctx.save_for_backward(torch.stack(states), x0, Kv, sg, m)
        return states[-1].clone()

    @staticmethod
    def backward(ctx, gout):
        states, x0, Kv, sg, m = ctx.saved_tensors
        sshape, mshape, sdt, mdt, it = ctx.meta
        B, _, H, W = gout.shape
        gout = _f32c(gout)
        g_sigma = torch.zeros(B, it, device=gout.device, dtype=torch.float32)
        g_mu = torch.zeros_like(g_sigma)
        g_state = torch.empty_like(gout)
        with torch.cuda.device(gout.device):
            _lib.check(_lib.lib().tfpnp_spi_admm_backward(
                ctx.solver.denoiser._grad_handle(gout.device), states.data_ptr(), x0.data_ptr(), Kv.data_ptr(), 1,
                sg.data_ptr(), m.data_ptr(), it, 1, B, H, W, it, gout.data_ptr(), g_sigma.data_ptr(), g_mu.data_ptr(),
                g_state.data_ptr(), torch.cuda.current_stream().cuda_stream), "tfpnp_spi_admm_backward")

        def widen(g, shape, dtype):
            full = torch.zeros(B, max(1, math.prod(shape[1:])), device=g.device, dtype=torch.float32)
            full[:, :it] = g
            return full.reshape(shape).to(dtype)

        return (None, None, g_state, None, None, widen(g_sigma, sshape, sdt), widen(g_mu, mshape, mdt), None)


# ---- factories, same names / error behaviour as the reference ---------------------------------
_csmri_map = {'admm': ADMMSolver_CSMRI, 'hqs': HQSSolver_CSMRI, 'pg': PGSolver_CSMRI, 'apg': APGSolver_CSMRI,
              'redadmm': REDADMMSolver_CSMRI}    # tasks/csmri/solver.py:253-270 ('amp' draws random numbers in the loop: not built)
_pr_map = {'iadmm': IADMMSolver_PR}          # tasks/pr/solver.py:115-128
_ct_map = {'iadmm': IADMMSolver_CT, 'pg': PGSolver_CT}          # tasks/ct/solver.py:90-103
_spi_map = {'admm_spi': ADMMSolver_SPI}      # tasks/spi/solver.py:54-66


def _create(opt, denoiser, table):
    print(f'[i] use solver: {opt.solver}')
    if opt.solver in table:
        return table[opt.solver](denoiser)
    raise NotImplementedError


def create_solver_csmri(opt, denoiser):
    return _create(opt, denoiser, _csmri_map)


def create_solver_pr(opt, denoiser):
    return _create(opt, denoiser, _pr_map)


def create_solver_ct(opt, denoiser):
    return _create(opt, denoiser, _ct_map)


def create_solver_spi(opt, denoiser):
    return _create(opt, denoiser, _spi_map)
