#!/usr/bin/env python
"""One small PR (256^2, 3 masks) and CT (64^2 and 40^2) update under compute-sanitizer: the round-2 kernels that exchange through shared
memory (pr256_*: half-warp FFT slots; ct: ray-split partial sums, back-projection windows).  fp32_simt denoiser, so the tensor-core
kernels (whose mbarrier / TMEM synchronisation racecheck does not model) stay out of the report.

    compute-sanitizer --tool memcheck|racecheck --error-exitcode 9 python tools/sanitize_updates.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, tfpnp_b200 as T
from oracle import synth

dev = torch.device("cuda:0")
den = T.UNetDenoiser2D(state_dict=synth.unet_state_dict(0, "default"), precision="fp32_simt")
cu = lambda d: {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}
with torch.no_grad():
    d = cu(synth.pr_batch(1, 256, 1, n_masks=3))
    s = T.IADMMSolver_PR(den); s.use_graph = False
    out = s((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"], d["tau"]))
    print("pr", float(out.abs().mean()))
    for n in (64, 40):
        g = torch.Generator().manual_seed(0)
        img = torch.rand(1, 1, n, n, generator=g).to(dev)
        y = T.radon_forward(img, 12)
        x = T.radon_backward(y, n, 12)
        print("radon", n, float(y.mean()), float(x.mean()))
    d = cu(synth.ct_batch(1, 64, 12, 1))
    s = T.IADMMSolver_CT(den); s.use_graph = False; s.opnorm_override = d["opnorm"]
    out = s((d["state"], (d["y0"], d["view"])), (d["sigma_d"], d["mu"], d["tau"]))
    print("ct", float(out.abs().mean()))
torch.cuda.synchronize()
