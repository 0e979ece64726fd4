"""Time the reverse mode at the north-star shape: one actor-update style call
    out = solver((state, aux), (sigma_d, mu)); out.backward(cotangent)
for CS-MRI ADMM, env_batch 48, 128x128, action_pack 5, with the convolutions of the VJP on CUDA cores (TFPNP_GRAD_TC=0) and on
the tensor cores (1: fp16, 2: split-fp16, 3: split-fp16 forward recompute + fp16 gradients).  Prints one JSON line per mode; also the agreement of modes 1/2 with mode 0.
Run on the GPU box:  python tools/grad_bench.py [--batch 48] [--size 128] [--iters 5]
"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tfpnp_b200 as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=48)
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    B, n, it = a.batch, a.size, a.iters
    g = torch.Generator().manual_seed(0)
    gt = torch.rand(B, 1, n, n, generator=g).to(dev)
    mask = T.radial_mask(n, max(4, n // 4)).to(dev)[None, None].expand(B, 1, n, n).contiguous()
    d = T.csmri_measure(gt, mask, 15 / 255)
    solver = T.ADMMSolver_CSMRI(T.UNetDenoiser2D(state_dict=T.random_unet_state_dict(0), precision="fp16"))
    solver.differentiable = True
    state = solver.reset(d)
    cot = torch.randn(state.shape, generator=g).to(dev)
    ref = None
    sg0 = (torch.rand(B, it, generator=g) * 70 / 255).to(dev)      # the SAME parameters for every mode
    mu0 = torch.rand(B, it, generator=g).to(dev)
    for mode in ("0", "1", "2", "3"):
        os.environ["TFPNP_GRAD_TC"] = mode
        times = []
        for r in range(a.reps + 1):
            sg = sg0.clone().requires_grad_(True)
            mu = mu0.clone().requires_grad_(True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = solver((state, (d["y0"], d["mask"])), (sg, mu))
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            gs, gm = torch.autograd.grad(out, (sg, mu), cot)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            if r:
                times.append((t1 - t0, t2 - t1))
        fwd = min(t[0] for t in times) * 1e3
        bwd = min(t[1] for t in times) * 1e3
        line = {"mode": int(mode), "batch": B, "size": n, "iters": it, "forward_ms": round(fwd, 2), "backward_ms": round(bwd, 2)}
        if ref is None:
            ref = (gs, gm)
        else:
            line["g_sigma_rel_l2_vs_fp32"] = ((gs - ref[0]).norm() / ref[0].norm()).item()
            line["g_mu_rel_l2_vs_fp32"] = ((gm - ref[1]).norm() / ref[1].norm()).item()
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
