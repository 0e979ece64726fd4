#!/usr/bin/env python
"""Reverse mode of every task once at its BASELINE per-GPU shape with action_pack = 5 (what one actor update differentiates):
csmri 48x128^2, pr 36x256^2 (4 masks), ct 8x256^2 (60 views), spi 48x128^2.  Checks finiteness, prints the forward / backward
wall time and the gradient norms as one JSON line per task.  TFPNP_GRAD_TC selects the VJP convolutions (default here: 3, the
split-fp16-forward / fp16-gradient tensor-core branch; 0 = CUDA cores, slow at 256^2).   python tools/grad_tasks.py [csmri|pr|ct|spi|all]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TFPNP_GRAD_TC", "3")
import torch  # noqa: E402
import tfpnp_b200 as T  # noqa: E402

dev = torch.device("cuda:0")
den = T.UNetDenoiser2D(state_dict=T.random_unet_state_dict(0), precision="fp16")
g = torch.Generator().manual_seed(0)
it = 5
only = sys.argv[1] if len(sys.argv) > 1 else "all"


def run(name, solver, state, aux, params):
    solver.differentiable = True
    ps = [p.to(dev).requires_grad_(True) for p in params]
    cot = torch.randn(state.shape, generator=g).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = solver((state, aux), tuple(ps))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    grads = torch.autograd.grad(out, ps, cot)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(json.dumps({"task": name, "batch": state.shape[0], "size": state.shape[2], "iters": it, "grad_tc": os.environ["TFPNP_GRAD_TC"],
                      "forward_ms": round((t1 - t0) * 1e3, 1), "backward_ms": round((t2 - t1) * 1e3, 1),
                      "finite": bool(all(torch.isfinite(x).all() for x in grads)),
                      "grad_norms": [round(x.norm().item(), 4) for x in grads]}), flush=True)


if only in ("all", "csmri"):
    B, n = 48, 128
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    d = T.csmri_measure(gt, T.radial_mask(n, 32).to(dev)[None, None], 15 / 255)
    s = T.ADMMSolver_CSMRI(den)
    run("csmri", s, s.reset(d), (d["y0"], d["mask"]), [torch.rand(B, it, generator=g) * 70 / 255, torch.rand(B, it, generator=g)])
if only in ("all", "spi"):
    B, n = 48, 128
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    d = T.spi_measure(gt, 6)
    s = T.ADMMSolver_SPI(den)
    K = torch.full((B, 1, n, n), 0.6, device=dev)
    run("spi", s, s.reset(d), (d["x0"], K), [(torch.rand(B, it, generator=g) * 55 + 15) / 255, torch.rand(B, it, generator=g) * 70 + 50])
if only in ("all", "ct"):
    B, n, views = 8, 256, 60
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    s = T.IADMMSolver_CT(den)
    s.opnorm_override = 200.0
    y0 = T.radon_forward(gt, views)
    x0 = T.radon_backward(y0, n, views) / 200.0 ** 2
    state = torch.cat((x0, x0.clone(), torch.zeros_like(x0)), 1)
    view = torch.full((B, 1, n, n), views / 120.0, device=dev)
    run("ct", s, state, (y0, view), [torch.rand(B, it, generator=g) * a for a in (70 / 255, 1.0, 2.0)])
if only in ("all", "pr"):
    B, n, M = 36, 256, 4
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    ph = torch.rand(B, M, n, n, generator=g).to(dev) * 6.283185307179586
    mask = torch.stack([torch.cos(ph), torch.sin(ph)], dim=-1)
    d = T.pr_measure(gt, mask, 27.0)
    s = T.IADMMSolver_PR(den)
    run("pr", s, s.reset(d), (d["y0"], d["mask"]), [torch.rand(B, it, generator=g) * a for a in (70 / 255, 1.0, 2.0)])
print("reverse-mode tasks done")
