"""Generate tests/golden/*.npz by running the UNMODIFIED reference (build container only).

    PYTHONPATH=/root/repo python -m oracle.make_golden

Each fixture stores the inputs, the reference's outputs and a checksum of the seeded UNet
weights used (the 47 MB weight set itself is regenerated from its seed by oracle/synth.py).
At generation time the CPU restatement (oracle/pnp_oracle.py) is asserted to agree with the
reference on the same inputs, so the fixtures pin both.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import pnp_oracle as O
from . import env_oracle as E
from . import refshim, synth

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def weight_checksum(sd) -> np.ndarray:
    """[sum, sum of squares, 8 strided samples] per tensor, fp64."""
    rows = []
    for k, v in sd.items():
        f = v.double().reshape(-1)
        idx = torch.linspace(0, f.numel() - 1, 8).long()
        rows.append(torch.cat([f.sum()[None], (f * f).sum()[None], f[idx]]))
    return torch.stack(rows).numpy()


def np_(d):
    return {k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in d.items()}


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


def close(a, b, tol=2e-6):
    err = (a - b).abs().max().item() / max(b.abs().max().item(), 1e-30)
    assert err <= tol, f"oracle deviates from the reference: {err}"
    return err


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    os.makedirs(OUT, exist_ok=True)
    refshim.install()
    weights = {("he", 0): synth.unet_state_dict(0, "he"), ("default", 0): synth.unet_state_dict(0, "default")}

    # reference state_dict layout check (tfpnp/pnp/denoiser/models/unet.py:34-47)
    from tfpnp.pnp.denoiser.models.unet import UNet
    ref_sd = UNet(2, 1).state_dict()
    assert [(k, tuple(v.shape)) for k, v in ref_sd.items()] == O.unet_param_shapes()

    with torch.no_grad():
        # 1. denoiser ------------------------------------------------------------------
        for (init, seed), sd in weights.items():
            g = torch.Generator().manual_seed(7)
            x = torch.rand(2, 1, 32, 32, generator=g)
            sigma = torch.tensor([10 / 255, 50 / 255])
            ref = refshim.reference_denoiser(sd)(x, sigma)
            e = close(O.denoise(sd, x, sigma), ref)
            save(f"denoiser_{init}", x=x.numpy(), sigma=sigma.numpy(), out=ref.numpy(),
                 wsum=weight_checksum(sd), init=init, seed=seed)
            print(f"  denoiser[{init}] oracle-vs-reference rel max err {e:.2e}")

        # 2. CS-MRI: small case (he) and BASELINE config 1 (B=4, 64x64, 6 iters, default init)
        for name, B, n, it, init in (("csmri_small", 2, 32, 3, "he"), ("csmri_cfg1", 4, 64, 6, "default")):
            sd = weights[(init, 0)]
            d = synth.csmri_batch(B, n, it)
            sol = refshim.reference_solver("csmri", sd)
            ref = sol((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))
            e = close(O.admm_csmri(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"]), ref)
            out_img = sol.get_output(ref)
            assert torch.equal(out_img, O.get_output(ref, True))
            save(name, **np_(d), out=ref.numpy(), wsum=weight_checksum(sd), init=init, seed=0,
                 psnr=O.psnr(out_img, d["gt"]).numpy())
            print(f"  {name} oracle-vs-reference rel max err {e:.2e}")
            # env metric (tfpnp/env/base.py:237-242)
            from tfpnp.env.base import torch_psnr
            assert torch.allclose(torch_psnr(out_img, d["gt"]), O.psnr(out_img, d["gt"]), rtol=1e-6)

        # 3. PR --------------------------------------------------------------------------
        sd = weights[("he", 0)]
        d = synth.pr_batch(2, 32, 3)
        sol = refshim.reference_solver("pr", sd)
        assert torch.equal(sol.reset({"x0": d["x0"]}), d["state"])
        ref = sol((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"], d["tau"]))
        e = close(O.iadmm_pr(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], d["tau"]), ref)
        save("pr_small", **np_(d), out=ref.numpy(), wsum=weight_checksum(sd), init="he", seed=0)
        print(f"  pr_small oracle-vs-reference rel max err {e:.2e}")

        # 4. SPI -------------------------------------------------------------------------
        d = synth.spi_batch(3, 32, 3)
        sol = refshim.reference_solver("spi", sd)
        ref = sol((d["state"], (d["x0"], d["K"])), (d["sigma_d"], d["mu"]))
        mine = O.admm_spi(sd, d["state"], d["x0"], d["K"], d["sigma_d"], d["mu"])
        e = close(mine, ref)
        save("spi_small", **np_(d), out=ref.numpy(), wsum=weight_checksum(sd), init="he", seed=0)
        print(f"  spi_small oracle-vs-reference rel max err {e:.2e}")

        # 4b. the SPI prox alone on a dense grid of operating points (transforms.py:404-439)
        from tfpnp.utils import transforms as T
        g = torch.Generator().manual_seed(3)
        zt = torch.rand(4, 1, 16, 16, generator=g) * 1.4 - 0.2
        Kv = torch.tensor([4.0, 6.0, 8.0, 6.0]).view(4, 1, 1, 1)
        K1 = torch.floor(torch.rand(4, 1, 16, 16, generator=g) * (Kv ** 2 + 1)).clamp(max=Kv ** 2)
        K1[:, :, :2] = 0
        mu = torch.tensor([50.0, 80.0, 120.0, 65.0]).view(4, 1, 1, 1)
        ref = T.spi_inverse(zt, K1, Kv, mu)
        assert torch.equal(O.spi_inverse(zt, K1, Kv, mu), ref)
        save("spi_prox", ztilde=zt.numpy(), K1=K1.numpy(), K=Kv.numpy(), mu=mu.numpy(), out=ref.numpy())

        # 5. transforms: centred FFT pair and CDP operators
        g = torch.Generator().manual_seed(5)
        x = torch.randn(2, 1, 32, 32, 2, generator=g)
        assert torch.equal(O.fft2c(x), T.fft2(x)) and torch.equal(O.ifft2c(x), T.ifft2(x))
        m = synth.pr_batch(2, 32, 1)["mask"]
        assert torch.equal(O.cdp_forward(x, m), T.cdp_forward(x, m))
        gg = torch.randn(2, 4, 32, 32, 2, generator=g)
        assert torch.equal(O.cdp_backward(gg, m), T.cdp_backward(gg, m))
        save("transforms", x=x.numpy(), fft2=T.fft2(x).numpy(), ifft2=T.ifft2(x).numpy(), mask=m.numpy(),
             cdp_fwd=T.cdp_forward(x, m).numpy(), g=gg.numpy(), cdp_bwd=T.cdp_backward(gg, m).numpy())

        # 5b. the other CS-MRI solvers of _solver_map (tasks/csmri/solver.py:60-201), driven through the reference classes
        d = synth.csmri_batch(2, 32, 3)
        g7 = torch.Generator().manual_seed(77)
        extra = {"tau": torch.rand(2, 3, generator=g7) * 1.5, "beta": torch.rand(2, 3, generator=g7) * 0.8,
                 "lamda": torch.rand(2, 3, generator=g7) * 0.5 + 0.05}
        x0 = d["x0"]
        rec = {}
        for name, fn, pk in (("hqs", O.hqs_csmri, ("sigma_d", "mu")), ("pg", O.pg_csmri, ("sigma_d", "tau")),
                             ("apg", O.apg_csmri, ("sigma_d", "tau", "beta")),
                             ("redadmm", O.redadmm_csmri, ("sigma_d", "mu", "lamda"))):
            sol = refshim.reference_solver("csmri_" + name, sd)
            state0 = sol.reset({"x0": x0})
            params = tuple({**d, **extra}[k] for k in pk)
            ref = sol((state0.clone(), (d["y0"], d["mask"])), params)
            e = close(fn(sd, state0.clone(), d["y0"], d["mask"], *params), ref)
            assert torch.equal(sol.get_output(ref), O.complex2real(torch.split(ref, ref.shape[1] // sol.num_var, dim=1)[0]))
            rec[name + "_state0"] = state0.numpy(); rec[name + "_out"] = ref.numpy()
            print(f"  csmri {name} oracle-vs-reference rel max err {e:.2e}")
        save("csmri_variants", **np_({k: d[k] for k in ("y0", "mask", "x0", "gt", "sigma_d", "mu")}), **np_(extra), **rec,
             wsum=weight_checksum(sd), init="he", seed=0)

        # 6. environment bookkeeping (tfpnp/env/base.py:121-191 + tasks/{csmri,spi}/env.py): a 3-step episode with
        #    images dropping out, driven through the UNMODIFIED reference env + solver classes
        env_fixture("csmri", weights[("he", 0)])
        env_fixture("spi", weights[("he", 0)])
    print("done")


def _load_ref_env(task):
    import importlib.util
    path = os.path.join(refshim.REFERENCE_ROOT, "tasks", task, "env.py")
    spec = importlib.util.spec_from_file_location(f"_ref_{task}_env", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return getattr(mod, {"csmri": "CSMRIEnv", "spi": "SPIEnv"}[task])


def env_fixture(task, sd, B=4, n=32, steps=3, pack=2):
    d = {"csmri": synth.csmri_batch, "spi": synth.spi_batch}[task](B, n, steps * pack)
    data = E.env_data(task, d)
    actions = E.episode_actions(task, d, B, steps, pack)
    ref_env = _load_ref_env(task)(None, refshim.reference_solver(task, sd), steps)
    ora = E.EnvOracle(task, sd, steps)
    clone = lambda x: {k: v.clone() for k, v in x.items()}
    ob_r = ref_env.reset(data=clone(data))
    ob_o = ora.reset(clone(data))
    rec = {}

    def check_ob(tag, r, o):
        pr, po = ref_env.get_policy_ob(r), ora.policy_ob(o)
        assert pr.shape == po.shape and (pr.shape[0] == 0 or close(po, pr) <= 2e-6), tag
        assert pr.shape[1] == ref_env.ob_base_dim + 3
        for k in o:
            assert torch.equal(torch.as_tensor(getattr(r, k)).float(), o[k].float()) or close(o[k].float(), torch.as_tensor(getattr(r, k)).float()) <= 2e-6, (tag, k)
        rec[tag + "_policy_ob"] = pr.numpy()
        rec[tag + "_variables"] = r.variables.numpy()

    check_ob("reset", ob_r, ob_o)
    for s, a in enumerate(actions):
        r = ref_env.step({k: v.clone() for k, v in a.items()})
        o = ora.step({k: v.clone() for k, v in a.items()})
        check_ob(f"step{s}_ob", r[0], o[0])
        check_ob(f"step{s}_masked", r[1], o[1])
        assert close(o[2], r[2], 1e-4) >= 0 and r[3] == o[3] and torch.equal(r[4]["done"], o[4]["done"])
        rec[f"step{s}_reward"] = r[2].numpy()
        rec[f"step{s}_all_done"] = np.asarray(r[3])
        rec[f"step{s}_done"] = r[4]["done"].numpy()
        for k, v in a.items():
            rec[f"step{s}_action_{k}"] = v.numpy()
        if r[3]:
            break
    rec["n_steps"] = np.asarray(s + 1)
    save(f"env_{task}", **{"data_" + k: v.numpy() for k, v in data.items()}, **rec, wsum=weight_checksum(sd), init="he",
         seed=0, max_episode_step=steps)
    print(f"  env_{task}: reference env vs oracle env agree over {s + 1} steps")


if __name__ == "__main__":
    main()
