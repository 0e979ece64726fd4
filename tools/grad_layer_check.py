"""Localise a reverse-mode failure on the GPU: run tfpnp_denoiser_vjp on the small fixture, copy back its workspace
(tfpnp_debug_grad_workspace) and compare it region by region -- the 27 kept activations, the pre-clamp output, the clamp-masked
cotangent, the per-level concatenation gradients, the final gradient buffers -- with the CPU emulation of the same code
(tests/grad_elem_host.cpp).  Prints one line per region; the first region that deviates names the kernel to look at.
    TFPNP_GRAD_TC=0|1|2 python tools/grad_layer_check.py
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tfpnp_b200 as T  # noqa: E402
from tfpnp_b200 import _lib  # noqa: E402
from tfpnp_b200.denoiser import flatten_state_dict  # noqa: E402
from conftest import load_golden, weights  # noqa: E402

NAMES = [f"act[{l}]" for l in range(27)] + ["in2", "pooled(tmp)", "upsampled(tmp)", "r (pre-clamp)", "g_r", "gA", "gB",
                                            "gcat[0]", "gcat[1]", "gcat[2]", "gcat[3]"]


def main():
    mode = int(os.environ.get("TFPNP_GRAD_TC", "0"))
    so = os.path.join(tempfile.mkdtemp(), "emu.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "grad_elem_host.cpp")], check=True)
    emu = C.CDLL(so)
    g = load_golden("grad_csmri_small")
    sd = weights("he")
    flat = flatten_state_dict(sd)
    x, s, go = g["den_x"].contiguous(), g["den_sigma"].contiguous(), g["den_gout"].contiguous()
    B, _, H, W = x.shape
    lay = (C.c_size_t * 39)()
    emu.emu_unet_vjp_layout(B, H, W, lay)
    total = lay[38]
    ws_cpu = torch.zeros(total)
    gx, gs = torch.zeros_like(x), torch.zeros(B)
    p = lambda t: C.c_void_p(t.data_ptr())
    emu.emu_unet_vjp_ws(p(flat), p(x), p(s), p(go), p(gx), p(gs), B, H, W, mode, p(ws_cpu))
    dev = torch.device("cuda:0")
    den = T.UNetDenoiser2D(state_dict=sd, precision="fp32_simt")
    mx, ms = den.vjp(x.to(dev), s.to(dev), go.to(dev))
    ws_gpu = torch.zeros(total)
    have = C.c_size_t()
    _lib.check(_lib.lib().tfpnp_debug_grad_workspace(den._grad_handle(dev), ws_gpu.data_ptr(), total, C.byref(have)), "workspace")
    print(f"mode {mode}: workspace {have.value} floats (emulation {total})")
    for k, name in enumerate(NAMES):
        a, b = ws_gpu[lay[k]:lay[k + 1]].double(), ws_cpu[lay[k]:lay[k + 1]].double()
        err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
        note = "   (temporaries hold the LAST level written: informative only)" if "tmp" in name or name in ("gA", "gB") else ""
        print(f"  {name:16s} rel max err {err:9.2e}{'  <-- deviates' if err > 1e-3 and not note else ''}{note}")
    print("  gx     ", ((mx.cpu() - gx).abs().max() / gx.abs().max()).item())
    print("  gsigma ", ((ms.cpu() - gs).abs().max() / gs.abs().max()).item())


if __name__ == "__main__":
    main()
