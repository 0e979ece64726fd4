#!/usr/bin/env python
"""How much of an inner iteration is per-launch fixed cost?  Times the graph-replayed CS-MRI loop (10 iterations) at several batch
sizes and fits  us_per_iteration = F + s * B:  F / 31 launches is the fixed cost per kernel inside the graph (with PDL)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfpnp_b200 as T

dev = torch.device("cuda:0")
n, iters = 128, 10
for prec in ("fp16", "fp16x3"):
    den = T.UNetDenoiser2D(state_dict=T.random_unet_state_dict(0), precision=prec)
    pts = []
    for B in (12, 24, 48, 96):
        solver = T.ADMMSolver_CSMRI(den)
        g = torch.Generator(dev).manual_seed(1)
        gt = torch.rand(B, 1, n, n, device=dev, generator=g)
        mask = T.radial_mask(n, n // 4, device=dev)[None, None].expand(B, 1, n, n).contiguous()
        m = T.csmri_measure(gt, mask, 15 / 255, generator=g)
        state = torch.cat((m["x0"], m["x0"].clone(), torch.zeros_like(m["x0"])), 1)
        sg = torch.rand(B, iters, device=dev) * 0.2
        mu = torch.rand(B, iters, device=dev)
        with torch.no_grad():
            for _ in range(3):
                solver((state, (m["y0"], m["mask"])), (sg, mu))
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); solver((state, (m["y0"], m["mask"])), (sg, mu)); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e3 / iters)
        pts.append((B, sorted(ts)[2]))
        del solver
    (b0, t0), (b1, t1) = pts[1], pts[2]
    s = (t1 - t0) / (b1 - b0)
    print(prec, " ".join(f"B={b}: {t:.0f} us/iter" for b, t in pts), f"| slope 24->48 {s:.2f} us/image, intercept {t0 - s * b0:.0f} us "
          f"= {(t0 - s * b0) / 31:.1f} us per launch; slope 48->96 {(pts[3][1] - t1) / 48:.2f}")
