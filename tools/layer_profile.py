#!/usr/bin/env python
"""Per-launch device times of one denoiser call, warm caches, back-to-back launches, no profiler attached
(tfpnp_denoiser_layer_profile): the in-situ companion of the ncu tables under profiles/.

    python tools/layer_profile.py [--precision fp16|fp16x3] [--batch 48] [--size 128] [--reps 20]

Prints one line per launch: time, share, achieved TFLOP/s of the conv layers (2*9*Cin*Cout*H*W*B), and the fraction of the
measured sustained tensor peak (MEASURED_PEAKS.json; fallback 1400)."""
import argparse
import ctypes as C
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tfpnp_b200 as T  # noqa: E402
from tfpnp_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = 1400.0
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk)).get("bf16_tflops_sustained", peak))
    den = T.UNetDenoiser2D(state_dict=T.random_unet_state_dict(0), precision=a.precision)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(a.batch, 1, a.size, a.size, generator=g).to(dev)
    sigma = (torch.rand(a.batch, generator=g) * 0.2).to(dev)
    out = torch.empty_like(x)
    cap = 128
    ms = (C.c_float * cap)()
    n = C.c_int(0)
    names = C.create_string_buffer(48 * cap)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().tfpnp_denoiser_layer_profile(den._handle(dev), x.data_ptr(), sigma.data_ptr(), out.data_ptr(),
                                                           a.batch, a.size, a.size, a.reps, ms, cap, C.byref(n), names,
                                                           torch.cuda.current_stream().cuda_stream), "layer_profile")
    tot = sum(ms[i] for i in range(n.value))
    print(f"# {a.precision}, B={a.batch}, {a.size}x{a.size}, mean of {a.reps} calls; events between eager launches (no PDL overlap); "
          f"peak = {peak:.0f} TFLOP/s sustained bf16")
    print(f"{'us':>8s} {'share':>6s} {'TFLOP/s':>8s} {'of peak':>7s}  launch")
    flops_tot = 0.0
    for i in range(n.value):
        name = names.raw[48 * i:48 * (i + 1)].split(b"\0")[0].decode()
        m = re.match(r"l(\d+) (\d+)->(\d+) @(\d+)", name)
        tf = ""
        if m:
            cin, cout, h = int(m.group(2)), int(m.group(3)), int(m.group(4))
            fl = 2.0 * 9 * cin * cout * h * h * a.batch
            flops_tot += fl
            t = fl / (ms[i] * 1e-3) / 1e12
            tf = f"{t:8.0f} {t / peak:7.2f}"
        print(f"{ms[i] * 1e3:8.1f} {100 * ms[i] / tot:5.1f}% {tf:>16s}  {name}")
    t = flops_tot / (tot * 1e-3) / 1e12
    print(f"{tot * 1e3:8.1f} us per call = {t:.0f} TFLOP/s algorithmic = {t / peak:.3f} of peak "
          f"({a.batch / (tot * 1e-3):.0f} image-iters/s denoiser-only)")


if __name__ == "__main__":
    main()
