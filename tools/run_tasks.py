#!/usr/bin/env python
"""Run each task's inner loop once at its BASELINE per-GPU shape (2 iterations) so a profiler can capture the
data-fidelity / ADMM-update kernels:  csmri 48x128^2, pr 36x256^2 (4 masks), ct 8x256^2 (60 views), spi 48x128^2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, tfpnp_b200 as T
from oracle import synth

dev = torch.device("cuda:0")
den = T.UNetDenoiser2D(state_dict=synth.unet_state_dict(0, "default"), precision="fp16")
g = torch.Generator().manual_seed(0)
it = 2
only = sys.argv[1] if len(sys.argv) > 1 else "all"
cu = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
with torch.no_grad():
  if only in ("all", "csmri"):
    d = cu(synth.csmri_batch(48, 128, it))
    s = T.ADMMSolver_CSMRI(den); s.use_graph = False
    for _ in range(2): s((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))
  if only in ("all", "pr"):
    d = cu(synth.pr_batch(36, 256, it))
    s = T.IADMMSolver_PR(den); s.use_graph = False
    for _ in range(2): s((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"], d["tau"]))
  if only in ("all", "spi"):
    d = cu(synth.spi_batch(48, 128, it))
    s = T.ADMMSolver_SPI(den); s.use_graph = False
    for _ in range(2): s((d["state"], (d["x0"], d["K"])), (d["sigma_d"], d["mu"]))
  if only in ("all", "ct"):
    # CT: measurements through the GPU operators (the CPU restatement of A at 256^2 x 60 views is slow)
    B, n, views = 8, 256, 60
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    s = T.IADMMSolver_CT(den); s.use_graph = False
    s.opnorm_override = 200.0          # any positive constant: the kernels' work does not depend on it
    y0 = T.radon_forward(gt, views)
    x0 = T.radon_backward(y0, n, views) / 200.0 ** 2
    state = torch.cat((x0, x0.clone(), torch.zeros_like(x0)), 1)
    view = torch.full((B, 1, n, n), views / 120.0, device=dev)
    p = [torch.rand(B, it, device=dev) * a for a in (70 / 255, 1.0, 2.0)]
    for _ in range(2): s((state, (y0, view)), tuple(p))
torch.cuda.synchronize()
print("tasks done")
