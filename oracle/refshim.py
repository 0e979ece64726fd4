"""Import the UNMODIFIED reference (Vandermode/TFPnP) on a modern PyTorch.
TEST INFRASTRUCTURE ONLY -- works only where ``/root/reference`` exists (the
build container); nothing that runs on the GPU box may call this.

The reference targets PyTorch <= 1.7 and calls ``torch.fft(x, 2, normalized=True)``
/ ``torch.ifft`` as functions (tfpnp/utils/transforms.py:4-5,82,101,300,318).
``install()`` replaces ``torch.fft`` by a callable module proxy and adds
``torch.ifft`` so those lines run unchanged; ``load_solver_module(task)`` loads
``tasks/<task>/solver.py`` by path (the four files share a module name).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import tempfile
import types

import torch

REFERENCE_ROOT = os.environ.get("TFPNP_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tfpnp"))


class _CallableFFT(types.ModuleType):
    def __init__(self, real):
        super().__init__("torch.fft")
        self.__dict__["_real"] = real

    def __getattr__(self, name):
        return getattr(self.__dict__["_real"], name)

    def __call__(self, x, signal_ndim, normalized=False):
        assert signal_ndim == 2
        c = torch.view_as_complex(x.contiguous())
        return torch.view_as_real(self._real.fft2(c, norm="ortho" if normalized else "backward"))


def _legacy_ifft(x, signal_ndim, normalized=False):
    assert signal_ndim == 2
    c = torch.view_as_complex(x.contiguous())
    import torch.fft as real  # the proxy forwards attribute access
    return torch.view_as_real(real.ifft2(c, norm="ortho" if normalized else "backward"))


def install():
    if not available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    if not isinstance(torch.fft, _CallableFFT):
        import torch.fft as real_fft
        proxy = _CallableFFT(real_fft)
        torch.fft = proxy
        torch.ifft = _legacy_ifft
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def load_solver_module(task: str):
    install()
    path = os.path.join(REFERENCE_ROOT, "tasks", task, "solver.py")
    spec = importlib.util.spec_from_file_location(f"_ref_{task}_solver", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_denoiser(state_dict):
    """UNetDenoiser2D (tfpnp/pnp/denoiser/base.py:7-32) loaded from a temp ckpt."""
    install()
    from tfpnp.pnp.denoiser import UNetDenoiser2D
    with tempfile.NamedTemporaryFile(suffix=".pt", delete=False) as f:
        torch.save(state_dict, f.name)
        path = f.name
    try:
        return UNetDenoiser2D(ckpt_path=path)
    finally:
        os.unlink(path)


def reference_solver(task: str, state_dict):
    cls = {"csmri": "ADMMSolver_CSMRI", "pr": "IADMMSolver_PR", "spi": "ADMMSolver_SPI",
           "csmri_hqs": "HQSSolver_CSMRI", "csmri_pg": "PGSolver_CSMRI", "csmri_apg": "APGSolver_CSMRI",
           "csmri_redadmm": "REDADMMSolver_CSMRI"}[task]
    mod = load_solver_module(task.split("_")[0])
    return getattr(mod, cls)(reference_denoiser(state_dict))
