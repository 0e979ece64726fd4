"""Generate tests/golden/grad_csmri_small.npz by differentiating the UNMODIFIED reference (build container only).

    PYTHONPATH=/root/repo python -m oracle.make_golden_grad

The reference's actor update back-propagates through ``ADMMSolver_CSMRI.forward`` with PyTorch autograd
(tfpnp/env/base.py:193-206, tfpnp/trainer/mddpg/trainer.py:173).  This script does the same on the seeded
``csmri_small`` case (2 images, 32x32, 3 iterations) with a seeded cotangent, records the gradients w.r.t.
``sigma_d`` / ``mu`` / the input state and the denoiser's vector-Jacobian product on its own, and asserts that both
restatements of oracle/grad_oracle.py (autograd through the oracle, and the hand-derived adjoint recursion the CUDA
path implements) agree with the reference.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import torch

from . import grad_oracle as G
from . import refshim, synth
from .make_golden import close, np_, save, weight_checksum


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    refshim.install()
    sd = synth.unet_state_dict(0, "he")

    # 1. denoiser VJP (tfpnp/pnp/denoiser/base.py:23-32 under autograd)
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 1, 32, 32, generator=g)
    sigma = torch.tensor([10 / 255, 50 / 255])
    gout = torch.randn(2, 1, 32, 32, generator=g)
    den = refshim.reference_denoiser(sd)
    xr, sr = x.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
    out = den(xr, sr)
    den_gx, den_gs = torch.autograd.grad(out, (xr, sr), gout)
    gx, gs = G.denoise_vjp_autograd(sd, x, sigma, gout)
    print(f"  denoiser vjp oracle-vs-reference: gx {close(gx, den_gx):.2e}  gsigma {close(gs, den_gs):.2e}")

    # 2. the solver: gradients of <forward(state, params), gout> w.r.t. sigma_d, mu and the state
    d = synth.csmri_batch(2, 32, 3)
    g = torch.Generator().manual_seed(13)
    gstate = torch.randn(d["state"].shape, generator=g)
    sol = refshim.reference_solver("csmri", sd)
    st = d["state"].clone().requires_grad_(True)
    sg = d["sigma_d"].clone().requires_grad_(True)
    mu = d["mu"].clone().requires_grad_(True)
    out = sol((st, (d["y0"], d["mask"])), (sg, mu))
    ref_gs, ref_gm, ref_gst = torch.autograd.grad(out, (sg, mu, st), gstate)
    a_gs, a_gm, a_gst = G.admm_csmri_vjp_autograd(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], gstate)
    print(f"  solver vjp (autograd through the oracle) vs reference: sigma_d {close(a_gs, ref_gs, 1e-5):.2e}  "
          f"mu {close(a_gm, ref_gm, 1e-5):.2e}  state {close(a_gst, ref_gst, 1e-5):.2e}")
    states = G.admm_csmri_trajectory(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"])
    m_gs, m_gm, m_gst = G.admm_csmri_vjp_manual(sd, states, d["y0"], d["mask"], d["sigma_d"], d["mu"], gstate)
    print(f"  solver vjp (adjoint recursion) vs reference:           sigma_d {close(m_gs, ref_gs, 1e-4):.2e}  "
          f"mu {close(m_gm, ref_gm, 1e-4):.2e}  state {close(m_gst, ref_gst, 1e-4):.2e}")

    save("grad_csmri_small", den_x=x.numpy(), den_sigma=sigma.numpy(), den_gout=gout.numpy(), den_gx=den_gx.numpy(),
         den_gsigma=den_gs.numpy(), **np_({k: d[k] for k in ("state", "y0", "mask", "sigma_d", "mu")}),
         gout=gstate.numpy(), g_sigma_d=ref_gs.numpy(), g_mu=ref_gm.numpy(), g_state=ref_gst.numpy(),
         wsum=weight_checksum(sd), init="he", seed=0)

    # 2b. SPI (tasks/spi/solver.py:17-51): the "differentiable binary search" of spi_inverse is a constant under autograd
    d = synth.spi_batch(3, 32, 3)
    g = torch.Generator().manual_seed(19)
    # a mid-episode state (x, z, u) instead of reset(): after reset the closed-form pixels (x0 == 0) sit below the clamp
    # (x + u - K0/mu < 0) and d/dmu would be identically zero
    d["state"] = torch.cat([torch.rand(3, 1, 32, 32, generator=g) * 1.2, torch.rand(3, 1, 32, 32, generator=g),
                            torch.rand(3, 1, 32, 32, generator=g) * 0.6], dim=1)
    gstate = torch.randn(d["state"].shape, generator=g)
    sol = refshim.reference_solver("spi", sd)
    st = d["state"].clone().requires_grad_(True)
    sg = d["sigma_d"].clone().requires_grad_(True)
    mu = d["mu"].clone().requires_grad_(True)
    out = sol((st, (d["x0"], d["K"])), (sg, mu))
    s_gs, s_gm, s_gst = torch.autograd.grad(out, (sg, mu, st), gstate)
    a = G.admm_spi_vjp_autograd(sd, d["state"], d["x0"], d["K"], d["sigma_d"], d["mu"], gstate)
    print(f"  SPI vjp (autograd through the oracle) vs reference: sigma_d {close(a[0], s_gs, 1e-5):.2e}  mu {close(a[1], s_gm, 1e-5):.2e}  "
          f"state {close(a[2], s_gst, 1e-5):.2e}")
    states = G.admm_spi_trajectory(sd, d["state"], d["x0"], d["K"], d["sigma_d"], d["mu"])
    m_ = G.admm_spi_vjp_manual(sd, states, d["x0"], d["K"], d["sigma_d"], d["mu"], gstate)
    print(f"  SPI vjp (adjoint recursion) vs reference:           sigma_d {close(m_[0], s_gs, 1e-4):.2e}  mu {close(m_[1], s_gm, 1e-4):.2e}  "
          f"state {close(m_[2], s_gst, 1e-4):.2e}   (closed-form pixels: {(d['x0'] == 0).float().mean():.3f})")
    save("grad_spi_small", **np_({k: d[k] for k in ("state", "x0", "K", "sigma_d", "mu")}), gout=gstate.numpy(),
         g_sigma_d=s_gs.numpy(), g_mu=s_gm.numpy(), g_state=s_gst.numpy(), wsum=weight_checksum(sd), init="he", seed=0)

    # 2c. PR (tasks/pr/solver.py:37-76)
    d = synth.pr_batch(2, 32, 3)
    g = torch.Generator().manual_seed(29)
    gstate = torch.randn(d["state"].shape, generator=g)
    sol = refshim.reference_solver("pr", sd)
    st = d["state"].clone().requires_grad_(True)
    ps = [d[k].clone().requires_grad_(True) for k in ("sigma_d", "mu", "tau")]
    out = sol((st, (d["y0"], d["mask"])), tuple(ps))
    p_ref = torch.autograd.grad(out, (*ps, st), gstate)
    a = G.iadmm_pr_vjp_autograd(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], d["tau"], gstate)
    print("  PR vjp (autograd through the oracle) vs reference: " + "  ".join(f"{n} {close(x, r, 1e-5):.2e}" for n, x, r in
          zip(("sigma_d", "mu", "tau", "state"), a, p_ref)))
    save("grad_pr_small", **np_({k: d[k] for k in ("state", "y0", "mask", "sigma_d", "mu", "tau")}), gout=gstate.numpy(),
         g_sigma_d=p_ref[0].numpy(), g_mu=p_ref[1].numpy(), g_tau=p_ref[2].numpy(), g_state=p_ref[3].numpy(),
         wsum=weight_checksum(sd), init="he", seed=0)

    # 2d. the other CS-MRI solvers of _solver_map (tasks/csmri/solver.py:60-201) under autograd
    from . import pnp_oracle as O
    d = synth.csmri_batch(2, 32, 3)
    g7 = torch.Generator().manual_seed(77)
    extra = {"tau": torch.rand(2, 3, generator=g7) * 1.5, "beta": torch.rand(2, 3, generator=g7) * 0.8,
             "lamda": torch.rand(2, 3, generator=g7) * 0.5 + 0.05}
    g = torch.Generator().manual_seed(37)
    rec = {}
    for name, fn, pk in (("hqs", O.hqs_csmri, ("sigma_d", "mu")), ("pg", O.pg_csmri, ("sigma_d", "tau")),
                         ("apg", O.apg_csmri, ("sigma_d", "tau", "beta")), ("redadmm", O.redadmm_csmri, ("sigma_d", "mu", "lamda"))):
        sol = refshim.reference_solver("csmri_" + name, sd)
        state0 = sol.reset({"x0": d["x0"]})
        cot = torch.randn(state0.shape, generator=g)
        ps = [{**d, **extra}[k].clone().requires_grad_(True) for k in pk]
        st = state0.clone().requires_grad_(True)
        out = sol((st, (d["y0"], d["mask"])), tuple(ps))
        ref = torch.autograd.grad(out, (*ps, st), cot)
        ps2 = [{**d, **extra}[k].clone().requires_grad_(True) for k in pk]
        st2 = state0.clone().requires_grad_(True)
        mine = torch.autograd.grad(fn(sd, st2, d["y0"], d["mask"], *ps2), (*ps2, st2), cot)
        print(f"  csmri {name} vjp (autograd through the oracle) vs reference: " + " ".join(f"{close(a, r, 1e-5):.1e}" for a, r in zip(mine, ref)))
        rec[name + "_state0"] = state0.numpy(); rec[name + "_gout"] = cot.numpy()
        for k, r in zip(pk + ("state",), ref):
            rec[f"{name}_g_{k}"] = r.numpy()
    save("grad_csmri_variants", **np_({k: d[k] for k in ("y0", "mask", "x0", "sigma_d", "mu")}), **np_(extra), **rec,
         wsum=weight_checksum(sd), init="he", seed=0)

    # 3. the call the trainer differentiates: ob2, reward = env.forward(ob, action) (tfpnp/env/base.py:193-206), loss through the
    #    next observation the critic reads (get_eval_ob) and through the PSNR reward (trainer.py:173-189)
    from . import env_oracle as E
    from .make_golden import _load_ref_env
    B, n, pack = 3, 32, 2
    d = synth.csmri_batch(B, n, pack)
    data = E.env_data("csmri", d)
    env = _load_ref_env("csmri")(None, refshim.reference_solver("csmri", sd), 3)
    ob = env.reset(data={k: v.clone() for k, v in data.items()})
    g = torch.Generator().manual_seed(17)
    sg = (d["sigma_d"][:, :pack].clone()).requires_grad_(True)
    mu = (d["mu"][:, :pack].clone()).requires_grad_(True)
    action = {"sigma_d": sg, "mu": mu, "idx_stop": torch.zeros(B, dtype=torch.long)}
    ob2, reward = env.forward(ob, action)
    eval2 = env.get_eval_ob(ob2)
    G1 = torch.randn(eval2.shape, generator=g)
    G2 = torch.randn(reward.shape, generator=g)
    loss = (eval2 * G1).sum() + (reward * G2).sum()
    e_gs, e_gm = torch.autograd.grad(loss, (sg, mu))
    print(f"  env.forward: reward {reward.detach().flatten().tolist()}, |d loss/d sigma_d| {e_gs.abs().max():.3e}, |d loss/d mu| {e_gm.abs().max():.3e}")
    save("grad_env_csmri", **{"data_" + k: v.numpy() for k, v in data.items()}, sigma_d=sg.detach().numpy(), mu=mu.detach().numpy(),
         G1=G1.numpy(), G2=G2.numpy(), eval_ob2=eval2.detach().numpy(), reward=reward.detach().numpy(), g_sigma_d=e_gs.numpy(),
         g_mu=e_gm.numpy(), wsum=weight_checksum(sd), init="he", seed=0, max_episode_step=3)

    # 3b. the same for SPI (tasks/spi/env.py): real-valued state, policy observation (variables, x0, K, T)
    d = synth.spi_batch(B, n, pack)
    data = E.env_data("spi", d)
    env = _load_ref_env("spi")(None, refshim.reference_solver("spi", sd), 3)
    ob = env.reset(data={k: v.clone() for k, v in data.items()})
    sg = (d["sigma_d"][:, :pack].clone()).requires_grad_(True)
    mu = (d["mu"][:, :pack].clone()).requires_grad_(True)
    ob2, reward = env.forward(ob, {"sigma_d": sg, "mu": mu, "idx_stop": torch.zeros(B, dtype=torch.long)})
    eval2 = env.get_eval_ob(ob2)
    G1 = torch.randn(eval2.shape, generator=g)
    G2 = torch.randn(reward.shape, generator=g)
    e_gs, e_gm = torch.autograd.grad((eval2 * G1).sum() + (reward * G2).sum(), (sg, mu), allow_unused=True)
    e_gm = torch.zeros_like(mu) if e_gm is None else e_gm
    print(f"  SPI env.forward: |d loss/d sigma_d| {e_gs.abs().max():.3e}, |d loss/d mu| {e_gm.abs().max():.3e}")
    save("grad_env_spi", **{"data_" + k: v.numpy() for k, v in data.items()}, sigma_d=sg.detach().numpy(), mu=mu.detach().numpy(),
         G1=G1.numpy(), G2=G2.numpy(), eval_ob2=eval2.detach().numpy(), reward=reward.detach().numpy(), g_sigma_d=e_gs.numpy(),
         g_mu=e_gm.numpy(), wsum=weight_checksum(sd), init="he", seed=0, max_episode_step=3)


if __name__ == "__main__":
    main()
