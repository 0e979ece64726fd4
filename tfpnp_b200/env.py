"""The caller side of the solver hot path: ``PnPEnv`` and the four task environments.

Mirrors tfpnp/env/base.py:43-242 (``PnPEnv.reset / step / forward / get_images / to``,
``torch_psnr``, ``_compute_metric / _compute_reward``) and tasks/{csmri,pr,ct,spi}/env.py
(``get_policy_ob``, ``get_eval_ob``, ``_get_attribute``, ``_build_next_ob``, ``_observation``,
``ob_base_dim``), so the RL plumbing (policy networks, trainer, evaluator) drives it unchanged.

What is different underneath: the per-step bookkeeping runs as a handful of fused
kernels of libtfpnp_b200 (include/tfpnp_b200.h, "environment bookkeeping") instead of the reference's
~40 indexing / cat / clone launches per step:

* ``_observation``       one multi-tensor row gather (``tfpnp_env_gather``)            base.py:176,188
* solver-result scatter  ``state['solver'][idx]`` and ``state['output'][idx]`` in one  base.py:171-172
* reward                 one PSNR kernel (``tfpnp_psnr``)                              base.py:230-242
* ``get_policy_ob``      one channel-packing kernel with the gather fused              tasks/*/env.py
* when every image is still active (``idx_left`` is the identity) the solver reads the state tensors in
  place -- no gather at all.

The only host synchronisation per step is the one the reference has too: the size of
``idx_left[idx_stop == 0]`` decides ``all_done`` (base.py:180-182).  No CPU fallback: the state must
live on a CUDA device.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .ops import torch_psnr


class Batch(dict):
    """Minimal attribute-style container standing in for tfpnp.data.batch.Batch (the reference's
    tianshou-derived class; only keyword construction, attribute/item access and row indexing are used
    on this path: tasks/csmri/env.py:14-57)."""

    def __init__(self, **kwargs):
        super().__init__(**kwargs)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    def __setattr__(self, key, value):
        self[key] = value

    def __getitem__(self, key):
        if isinstance(key, str):
            return dict.__getitem__(self, key)
        return Batch(**{k: v[key] for k, v in self.items()})

    def to(self, device):
        return Batch(**{k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.items()})

    @property
    def shape(self):
        for v in self.values():
            if isinstance(v, torch.Tensor):
                return [v.shape[0]]
        return []


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


def gather_rows(tensors, idx):
    """[t[idx] for t in tensors] in ONE launch.  ``idx`` int64 CUDA vector or None (identity copy)."""
    tensors = [t.contiguous() for t in tensors]
    dev = tensors[0].device
    if not dev.type == "cuda":
        raise RuntimeError("tfpnp_b200.env runs on CUDA tensors only; there is no CPU fallback")
    n = int(idx.shape[0]) if idx is not None else int(tensors[0].shape[0])
    outs = [torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for t in tensors]
    items = (_lib.GatherItem * len(tensors))()
    for it, t, o in zip(items, tensors, outs):
        row = t[0].numel() * t.element_size() if t.shape[0] else 0
        it.src, it.dst, it.row_bytes = t.data_ptr(), o.data_ptr(), max(row, 1)
    if n:
        with torch.cuda.device(dev):
            for lo in range(0, len(tensors), 12):
                k = min(12, len(tensors) - lo)
                sub = (_lib.GatherItem * k)(*items[lo:lo + k])
                _lib.check(_lib.lib().tfpnp_env_gather(sub, k, idx.data_ptr() if idx is not None else None, n,
                                                       _stream(dev)), "tfpnp_env_gather")
    return outs


class Env:
    """tfpnp/env/base.py:10-36."""

    def reset(self):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError


class DifferentiableEnv(Env):
    """tfpnp/env/base.py:38-40."""

    def forward(self, state, action):
        raise NotImplementedError


class PnPEnv(DifferentiableEnv):
    """tfpnp/env/base.py:43-234."""

    # (key, kind) list describing get_policy_ob's channels, filled by the task subclasses:
    #   'real'     : real plane(s)            [B,C,H,W]      -> C channels
    #   'c2real'   : complex2real             [B,C,H,W,2]    -> C channels (re)      transforms.py:16-17
    #   'c2chan'   : complex2channel          [B,C,H,W,2]    -> 2C channels (re,im)  transforms.py:20-26
    _policy_channels = ()
    _ob_keys = ()                 # state keys an observation carries (besides 'variables')
    _complex_state = False

    def __init__(self, data_loader, solver, max_episode_step, data_transform=None):
        super().__init__()
        self.data_loader = data_loader
        self.data_iterator = iter(data_loader) if data_loader is not None else None
        self.device = torch.device('cpu')
        self.data_transform = data_transform
        self.solver = solver
        self.max_episode_step = max_episode_step
        self.cur_step = 0
        self.state = None
        self.last_metric = 0
        self.metric_fn = torch_psnr
        self.idx_left = None
        self._all_active = True

    # ---- abstract (tasks/*/env.py) ------------------------------------------------------------
    def get_policy_ob(self, ob):
        raise NotImplementedError

    def get_eval_ob(self, ob):
        return self.get_policy_ob(ob)

    def _get_attribute(self, ob, key):
        raise NotImplementedError

    def _build_next_ob(self, ob, solver_state):
        raise NotImplementedError

    # ---- basic API (base.py:121-223) ----------------------------------------------------------
    def reset(self, data=None):
        self.cur_step = 0
        if data is None:                                                  # base.py:125-130
            try:
                data = next(self.data_iterator)
            except StopIteration:
                self.data_iterator = iter(self.data_loader)
                data = next(self.data_iterator)
        if self.data_transform is not None:
            data = self.data_transform(data)
        if self.device.type != "cuda":
            raise RuntimeError("tfpnp_b200.PnPEnv needs .to(torch.device('cuda', i)) first: the environment state "
                               "lives on the GPU; there is no CPU fallback")
        data = {k: (v.to(self.device) if isinstance(v, torch.Tensor) else v) for k, v in data.items()}
        solver_state = self.solver.reset(data)                            # base.py:141
        data['solver'] = solver_state.contiguous()
        data['output'] = data['output'].contiguous().float()
        B, _, W, H = data['gt'].shape
        data['T'] = torch.full([B, 1, W, H], self.cur_step / self.max_episode_step, dtype=torch.float32,
                               device=self.device)                       # base.py:147-149
        self.state = data
        self.idx_left = torch.arange(0, B, device=self.device)
        self._all_active = True
        self.last_metric = self._compute_metric()
        return self._observation()

    def step(self, action):
        self.cur_step += 1
        idx = None if self._all_active else self.idx_left
        with torch.no_grad():
            aux = self.solver.filter_aux_inputs(self.state)
            if idx is None:                                               # every image active: read in place
                inputs = (self.state['solver'], tuple(aux))
            else:
                g = gather_rows([self.state['solver']] + list(aux), idx)  # base.py:162-166
                inputs = (g[0], tuple(g[1:]))
            parameters = self.solver.filter_hyperparameter(action)
            solver_state = self.solver(inputs, parameters)                # base.py:168

        self.state['T'].fill_(self.cur_step / self.max_episode_step)      # base.py:170
        self._scatter_state(solver_state, idx)                            # base.py:171-172
        reward = self._compute_reward()
        ob = self._observation()

        idx_stop = action['idx_stop']                                     # base.py:179-186
        self.idx_left = self.idx_left[idx_stop == 0]
        n_left = int(self.idx_left.shape[0])                              # the one host sync of a step
        self._all_active = self._all_active and n_left == self.state['gt'].shape[0]
        all_done = n_left == 0
        done = idx_stop.detach()
        if self.cur_step == self.max_episode_step:
            all_done = True
            done = torch.ones_like(idx_stop)
        ob_masked = self._observation()
        return ob, ob_masked, reward, all_done, {'done': done}

    def forward(self, ob, action):                                        # base.py:193-206
        output = self._get_attribute(ob, 'output')
        gt = self._get_attribute(ob, 'gt')
        inputs = self._get_attribute(ob, 'solver_input')
        parameters = self.solver.filter_hyperparameter(action)
        solver_state = self.solver(inputs, parameters)    # under autograd: opt-in, solver.differentiable = True
        output2 = self.solver.get_output(solver_state)
        reward = self.metric_fn(output2, gt) - self.metric_fn(output, gt)
        return self._build_next_ob(ob, solver_state), reward

    def get_images(self, ob, pre_process=None):                           # base.py:208-213
        if pre_process is None:
            def pre_process(img):                                         # torch2img255, tfpnp/utils/misc.py
                return (img.clamp(0, 1) * 255).round().to(torch.uint8).cpu().numpy()
        return (pre_process(self._get_attribute(ob, 'input')), pre_process(self._get_attribute(ob, 'output')),
                pre_process(self._get_attribute(ob, 'gt')))

    def to(self, device):                                                 # base.py:215-219
        if not isinstance(device, torch.device):
            raise TypeError('device must be torch.device, but got {}'.format(type(device)))
        self.device = device
        return self

    # ---- private -------------------------------------------------------------------------------
    def _compute_metric(self):                                            # base.py:225-228
        return self.metric_fn(self.state['output'], self.state['gt'])

    def _compute_reward(self):                                            # base.py:230-234
        metric = self._compute_metric()
        reward = metric - self.last_metric
        self.last_metric = metric
        return reward

    def _scatter_state(self, solver_state, idx):
        s = solver_state.contiguous()
        st, out = self.state['solver'], self.state['output']
        HW = out[0].numel()
        n = s.shape[0]
        # the reference's index_put works for any solver.num_var (3 ADMM, 2 HQS / APG, 1 PG); rows must agree
        if tuple(s.shape[1:]) != tuple(st.shape[1:]) or s.dtype != torch.float32 or st.dtype != torch.float32:
            raise ValueError(f"solver returned state rows {tuple(s.shape[1:])} {s.dtype}, the environment holds "
                             f"{tuple(st.shape[1:])} {st.dtype}")
        num_var = int(s.shape[1])
        if s[0].numel() != num_var * HW * (2 if self._complex_state else 1):
            raise ValueError(f"state row of {s[0].numel()} floats is not num_var={num_var} x HW={HW} x "
                             f"{'2 (complex)' if self._complex_state else '1'}")
        if idx is not None and int(idx.shape[0]) != n:
            raise ValueError(f"{n} solver rows for {int(idx.shape[0])} active images")
        with torch.cuda.device(s.device):
            _lib.check(_lib.lib().tfpnp_env_scatter_state(
                s.data_ptr(), idx.data_ptr() if idx is not None else None, n, st.data_ptr(), out.data_ptr(), HW,
                1 if self._complex_state else 0, num_var, _stream(s.device)), "tfpnp_env_scatter_state")

    def _observation(self):
        """tasks/*/env.py `_observation`: every tensor of the state restricted to idx_left, one launch."""
        keys = list(self._ob_keys)
        src = [self.state['solver']] + [self.state[k] for k in keys]
        out = gather_rows(src, None if self._all_active else self.idx_left)
        ob = Batch(variables=out[0])
        for k, v in zip(keys, out[1:]):
            ob[k] = v.float() if k == 'mask' and v.dtype == torch.bool else v      # mask=...float() (csmri/env.py:55)
        return ob

    def _pack_policy_ob(self, ob):
        """One launch for the `torch.cat([...], 1)` of tasks/*/env.py get_policy_ob."""
        chans = []
        keep = []
        first = None
        for key, kind in self._policy_channels:
            t = ob[key]
            if not t.is_cuda:
                raise RuntimeError("tfpnp_b200.env runs on CUDA tensors only; there is no CPU fallback")
            if t.dtype == torch.bool:
                t = t.contiguous().view(torch.uint8)
                dtype = 1
            else:
                t = t.contiguous() if t.dtype == torch.float32 else t.float().contiguous()
                dtype = 0
            keep.append(t)
            first = t if first is None else first
            Cn = t.shape[1]
            if kind == 'real':
                HW = t.shape[2] * t.shape[3]
                for c in range(Cn):
                    chans.append((t.data_ptr(), Cn * HW, c * HW, 1, dtype))
            elif kind == 'c2real':
                HW = t.shape[2] * t.shape[3]
                for c in range(Cn):
                    chans.append((t.data_ptr(), Cn * HW * 2, c * HW * 2, 2, dtype))
            elif kind == 'c2chan':
                HW = t.shape[2] * t.shape[3]
                for c in range(Cn):
                    for part in range(2):
                        chans.append((t.data_ptr(), Cn * HW * 2, c * HW * 2 + part, 2, dtype))
            else:
                raise ValueError(kind)
        B, H, W = first.shape[0], first.shape[2], first.shape[3]
        dst = torch.empty(B, len(chans), H, W, dtype=torch.float32, device=first.device)
        arr = (_lib.ObChannel * len(chans))()
        for a, (ptr, istr, off, ps, dt) in zip(arr, chans):
            a.src, a.img_stride, a.offset, a.pix_stride, a.dtype = ptr, istr, off, ps, dt
        if B:
            with torch.cuda.device(first.device):
                _lib.check(_lib.lib().tfpnp_env_policy_ob(arr, len(chans), None, B, H * W, dst.data_ptr(),
                                                          _stream(first.device)), "tfpnp_env_policy_ob")
        v = ob['variables']
        if torch.is_grad_enabled() and v.requires_grad:
            # env.forward under autograd (trainer.py:173-187): the critic reads the solver variables of the next
            # observation through these channels, so they must carry a gradient.  The variables are the leading
            # channels ('variables' is first in every _policy_channels); the values are the kernel's.
            key, kind = self._policy_channels[0]
            assert key == 'variables'
            nv = v.shape[1]
            return _PackObFn.apply(v, dst, nv, kind == 'c2real')
        return dst


class _PackObFn(torch.autograd.Function):
    """Identity on the packed observation that routes d/d(ob[:, :nv]) back to the solver variables (pure data movement:
    'c2real' takes the real part of a complex variable, tfpnp/utils/transforms.py:16-17)."""

    @staticmethod
    def forward(ctx, variables, packed, nv, complex_state):
        ctx.nv, ctx.complex_state, ctx.vshape = nv, complex_state, variables.shape
        return packed.clone()

    @staticmethod
    def backward(ctx, g):
        gv = torch.zeros(ctx.vshape, device=g.device, dtype=g.dtype)
        if ctx.complex_state:
            gv[..., 0] = g[:, :ctx.nv]
        else:
            gv.copy_(g[:, :ctx.nv])
        return gv, None, None, None


class CSMRIEnv(PnPEnv):
    """tasks/csmri/env.py:7-57."""
    ob_base_dim = 6
    _complex_state = True
    _ob_keys = ('gt', 'y0', 'ATy0', 'mask', 'sigma_n', 'T')
    _policy_channels = (('variables', 'c2real'), ('y0', 'c2chan'), ('ATy0', 'c2real'), ('mask', 'real'),
                        ('T', 'real'), ('sigma_n', 'c2real'))

    def __init__(self, data_loader, solver, max_episode_step):
        super().__init__(data_loader, solver, max_episode_step)

    def get_policy_ob(self, ob):
        return self._pack_policy_ob(ob)

    def _get_attribute(self, ob, key):
        if key == 'gt':
            return ob.gt
        elif key == 'output':
            return self.solver.get_output(ob.variables)
        elif key == 'input':
            return ob.ATy0
        elif key == 'solver_input':
            return ob.variables, (ob.y0, ob.mask.bool())
        raise NotImplementedError('key is not supported, ' + str(key))

    def _build_next_ob(self, ob, solver_state):
        return Batch(gt=ob.gt, y0=ob.y0, ATy0=ob.ATy0, variables=solver_state, mask=ob.mask, sigma_n=ob.sigma_n,
                     T=ob.T + 1 / self.max_episode_step)


class PREnv(PnPEnv):
    """tasks/pr/env.py:7-56."""
    ob_base_dim = 14
    _complex_state = True
    _ob_keys = ('gt', 'y0', 'x0', 'mask', 'sigma_n', 'T')
    _policy_channels = (('variables', 'c2real'), ('y0', 'real'), ('mask', 'c2chan'), ('T', 'real'), ('sigma_n', 'real'))

    def __init__(self, data_loader, solver, max_episode_step):
        super().__init__(data_loader, solver, max_episode_step)

    def get_policy_ob(self, ob):
        return self._pack_policy_ob(ob)

    def _get_attribute(self, ob, key):
        if key == 'gt':
            return ob.gt
        elif key == 'output':
            return self.solver.get_output(ob.variables)
        elif key == 'input':
            return ob.x0
        elif key == 'solver_input':
            return (ob.variables, (ob.y0, ob.mask))
        raise NotImplementedError('key is not supported, ' + str(key))

    def _build_next_ob(self, ob, solver_state):
        return Batch(gt=ob.gt, y0=ob.y0, x0=ob.x0, variables=solver_state, mask=ob.mask, sigma_n=ob.sigma_n,
                     T=ob.T + 1 / self.max_episode_step)


class CTEnv(PnPEnv):
    """tasks/ct/env.py:6-54."""
    ob_base_dim = 4
    _ob_keys = ('gt', 'y0', 'ATy0', 'view', 'sigma_n', 'T')
    _policy_channels = (('variables', 'real'), ('ATy0', 'real'), ('view', 'real'), ('T', 'real'), ('sigma_n', 'real'))

    def __init__(self, data_loader, solver, max_episode_step, data_transform=None):
        super().__init__(data_loader, solver, max_episode_step, data_transform)

    def get_policy_ob(self, ob):
        return self._pack_policy_ob(ob)

    def _get_attribute(self, ob, key):
        if key == 'gt':
            return ob.gt
        elif key == 'output':
            return self.solver.get_output(ob.variables)
        elif key == 'input':
            return ob.ATy0
        elif key == 'solver_input':
            return (ob.variables, (ob.y0, ob.view))
        raise NotImplementedError('key is not supported, ' + str(key))

    def _build_next_ob(self, ob, solver_state):
        return Batch(gt=ob.gt, y0=ob.y0, ATy0=ob.ATy0, variables=solver_state, view=ob.view, sigma_n=ob.sigma_n,
                     T=ob.T + 1 / self.max_episode_step)


class SPIEnv(PnPEnv):
    """tasks/spi/env.py:6-50."""
    ob_base_dim = 3
    _ob_keys = ('gt', 'x0', 'K', 'T')
    _policy_channels = (('variables', 'real'), ('x0', 'real'), ('K', 'real'), ('T', 'real'))

    def __init__(self, data_loader, solver, max_episode_step):
        super().__init__(data_loader, solver, max_episode_step)

    def get_policy_ob(self, ob):
        return self._pack_policy_ob(ob)

    def _get_attribute(self, ob, key):
        if key == 'gt':
            return ob.gt
        elif key == 'output':
            return self.solver.get_output(ob.variables)
        elif key == 'input':
            return ob.x0
        elif key == 'solver_input':
            return (ob.variables, (ob.x0, ob.K))
        raise NotImplementedError('key is not supported, ' + str(key))

    def _build_next_ob(self, ob, solver_state):
        return Batch(gt=ob.gt, x0=ob.x0, variables=solver_state, K=ob.K, T=ob.T + 1 / self.max_episode_step)
