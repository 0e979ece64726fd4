#!/usr/bin/env python
"""Does splitting the batch over concurrent streams hide the per-launch fixed cost and the under-filled deep levels?
Runs the graph-replayed CS-MRI loop (30 iterations, 128^2) for 48 images as 1 x 48, 2 x 24, 3 x 16 and 4 x 12 on as many streams
(one solver + denoiser instance per stream: separate workspaces) and prints image-iterations/s."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfpnp_b200 as T

dev = torch.device("cuda:0")
n, iters, B = 128, 30, 48
for prec in ("fp16", "fp16x3"):
    for parts in (1, 2, 3, 4):
        b = B // parts
        streams = [torch.cuda.Stream() for _ in range(parts)]
        jobs = []
        for p in range(parts):
            den = T.UNetDenoiser2D(state_dict=T.random_unet_state_dict(0), precision=prec)
            solver = T.ADMMSolver_CSMRI(den)
            g = torch.Generator(dev).manual_seed(1 + p)
            gt = torch.rand(b, 1, n, n, device=dev, generator=g)
            mask = T.radial_mask(n, n // 4, device=dev)[None, None].expand(b, 1, n, n).contiguous()
            m = T.csmri_measure(gt, mask, 15 / 255, generator=g)
            state = torch.cat((m["x0"], m["x0"].clone(), torch.zeros_like(m["x0"])), 1)
            sg = torch.rand(b, iters, device=dev) * 0.2
            mu = torch.rand(b, iters, device=dev)
            jobs.append((solver, (state, (m["y0"], m["mask"])), (sg, mu)))
        torch.cuda.synchronize()
        def run():
            for st, (solver, a, q) in zip(streams, jobs):
                with torch.cuda.stream(st):
                    solver(a, q)
        with torch.no_grad():
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0 = torch.cuda.Event(enable_timing=True); e0.record()
                for st in streams: st.wait_event(e0)
                run()
                ends = []
                for st in streams:
                    e = torch.cuda.Event(enable_timing=True); e.record(st); ends.append(e)
                torch.cuda.synchronize()
                ts.append(max(e0.elapsed_time(e) for e in ends))
        t = sorted(ts)[2]
        print(f"{prec} {parts} x {b}: {t:.2f} ms -> {B * iters / t * 1e3:.0f} image-iterations/s", flush=True)
        del jobs
