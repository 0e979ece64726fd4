"""ctypes binding of libtfpnp_b200.so (the C ABI declared in include/tfpnp_b200.h).

The library is built in-tree by ``tfpnp_b200/csrc/Makefile`` (see ``__graft_entry__.build``).
There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is
raised -- the product path never silently degrades to PyTorch/CPU code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtfpnp_b200.so")

TASK_CSMRI, TASK_PR, TASK_CT, TASK_SPI = 0, 1, 2, 3
ALGO_HQS, ALGO_PG, ALGO_APG, ALGO_REDADMM = 1, 2, 3, 4
PREC_FP16, PREC_FP16X3, PREC_FP32_SIMT = 0, 1, 2
PRECISIONS = {"fp16": PREC_FP16, "fp16x3": PREC_FP16X3, "fp32_simt": PREC_FP32_SIMT}


class SolverConfig(C.Structure):
    _fields_ = [("task", C.c_int), ("H", C.c_int), ("W", C.c_int), ("n_masks", C.c_int),
                ("views", C.c_int), ("opnorm", C.c_float), ("ct_cos", C.POINTER(C.c_float)),
                ("ct_sin", C.POINTER(C.c_float)), ("use_graph", C.c_int)]


class GatherItem(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("row_bytes", C.c_int64)]


class ObChannel(C.Structure):
    _fields_ = [("src", C.c_void_p), ("img_stride", C.c_int64), ("offset", C.c_int64),
                ("pix_stride", C.c_int32), ("dtype", C.c_int32)]


# every symbol include/tfpnp_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tfpnp_version": (C.c_int, []),
    "tfpnp_last_error": (C.c_char_p, []),
    "tfpnp_release_cached_scratch": (C.c_int, []),
    "tfpnp_denoiser_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tfpnp_ircnn_create": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tfpnp_denoiser_destroy": (C.c_int, [C.c_void_p]),
    "tfpnp_denoiser_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                         C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tfpnp_denoiser_vjp": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tfpnp_debug_grad_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "tfpnp_denoiser_layer_profile": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int), C.c_char_p,
                                               C.c_void_p]),
    "tfpnp_conv3x3_nhwc": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tfpnp_solver_create": (C.c_int, [C.POINTER(SolverConfig), C.c_void_p, C.POINTER(C.c_void_p)]),
    "tfpnp_solver_destroy": (C.c_int, [C.c_void_p]),
    "tfpnp_solver_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tfpnp_solver_last_launch_count": (C.c_int64, [C.c_void_p]),
    "tfpnp_csmri_variant_create": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "tfpnp_csmri_variant_destroy": (C.c_int, [C.c_void_p]),
    "tfpnp_csmri_variant_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "tfpnp_csmri_admm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfpnp_spi_admm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfpnp_ct_iadmm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfpnp_pr_iadmm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfpnp_csmri_variant_backward": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "tfpnp_radon_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "tfpnp_radon_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "tfpnp_fft2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "tfpnp_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "tfpnp_psnr_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    "tfpnp_env_gather": (C.c_int, [C.POINTER(GatherItem), C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "tfpnp_env_scatter_state": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_int, C.c_int, C.c_void_p]),
    "tfpnp_env_policy_ob": (C.c_int, [C.POINTER(ObChannel), C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_void_p,
                                      C.c_void_p]),
    "tfpnp_comm_unique_id": (C.c_int, [C.c_void_p, C.c_size_t]),
    "tfpnp_comm_init": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "tfpnp_comm_destroy": (C.c_int, [C.c_void_p]),
    "tfpnp_comm_allgather_psnr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "tfpnp_solver_set_profiling": (C.c_int, [C.c_void_p, C.c_int]),
    "tfpnp_solver_get_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile libtfpnp_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libtfpnp_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: the CUDA library has not been built "
                "(run `python -c 'import __graft_entry__ as g; g.build()'` or `make -C tfpnp_b200/csrc`). "
                "tfpnp_b200 has no CPU/PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status: int, what: str):
    if status != 0:
        msg = lib().tfpnp_last_error()
        raise RuntimeError(f"{what} failed (status {status}): {msg.decode() if msg else '?'}")


def release_cached_scratch():
    """Free the per-(device, stream) pool of device blocks the reverse-mode entry points draw their scratch from
    (the counterpart of torch.cuda.empty_cache() for this library)."""
    check(lib().tfpnp_release_cached_scratch(), "tfpnp_release_cached_scratch")
