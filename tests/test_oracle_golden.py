"""CPU tests: the oracle restatement against the golden vectors produced by the unmodified
reference (oracle/make_golden.py), plus structural pins for the CT operators whose reference
arithmetic (torch_radon) is absent."""
import math
import os

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err, weights
from oracle import pnp_oracle as O
from oracle import refshim, synth


def _wsum(sd):
    from oracle.make_golden import weight_checksum
    return weight_checksum(sd)


@pytest.mark.parametrize("init", ["he", "default"])
def test_seeded_weights_match_fixture(init):
    g = load_golden(f"denoiser_{init}")
    np.testing.assert_allclose(_wsum(weights(init)), g["wsum"].numpy(), rtol=0, atol=0)


def test_state_dict_layout():
    shapes = O.unet_param_shapes()
    assert len(shapes) == 56
    assert sum(int(np.prod(s)) for _, s in shapes) == 11773857


@pytest.mark.parametrize("init", ["he", "default"])
def test_denoiser_golden(init):
    g = load_golden(f"denoiser_{init}")
    out = O.denoise(weights(init), g["x"], g["sigma"])
    assert torch.equal(out, g["out"])


@pytest.mark.parametrize("name", ["csmri_small", "csmri_cfg1"])
def test_csmri_golden(name):
    g = load_golden(name)
    sd = weights(str(g["init"]))
    out = O.admm_csmri(sd, g["state"], g["y0"], g["mask"], g["sigma_d"], g["mu"])
    assert rel_err(out, g["out"])[1] <= 2e-6
    p = O.psnr(O.get_output(out, True), g["gt"])
    assert torch.allclose(p, g["psnr"], rtol=1e-5)


def test_pr_golden():
    g = load_golden("pr_small")
    out = O.iadmm_pr(weights("he"), g["state"], g["y0"], g["mask"], g["sigma_d"], g["mu"], g["tau"])
    assert rel_err(out, g["out"])[1] <= 2e-6
    assert torch.equal(O.pr_reset(g["x0"]), g["state"])


def test_spi_golden():
    g = load_golden("spi_small")
    out = O.admm_spi(weights("he"), g["state"], g["x0"], g["K"], g["sigma_d"], g["mu"])
    assert rel_err(out, g["out"])[1] <= 2e-6


def test_spi_prox_golden_bit_exact():
    g = load_golden("spi_prox")
    out = O.spi_inverse(g["ztilde"], g["K1"], g["K"], g["mu"])
    assert torch.equal(out, g["out"])
    # closed form where K1 == 0 (transforms.py:415)
    z0 = torch.clamp(g["ztilde"] - (g["K"] ** 2) / g["mu"], 0, 1)
    assert torch.equal(out[:, :, :2], z0[:, :, :2])


def test_transforms_golden():
    g = load_golden("transforms")
    assert torch.equal(O.fft2c(g["x"]), g["fft2"])
    assert torch.equal(O.ifft2c(g["x"]), g["ifft2"])
    assert torch.equal(O.cdp_forward(g["x"], g["mask"]), g["cdp_fwd"])
    assert torch.equal(O.cdp_backward(g["g"], g["mask"]), g["cdp_bwd"])
    # unitarity of the centred pair
    assert rel_err(O.ifft2c(O.fft2c(g["x"])), g["x"])[1] < 1e-6


def test_fp64_oracle_budget():
    """fp32 vs fp64 run of the same loop: the error floor any fp32 implementation has."""
    g = load_golden("csmri_small")
    sd = weights("he")
    sd64 = {k: v.double() for k, v in sd.items()}
    o32 = O.admm_csmri(sd, g["state"], g["y0"], g["mask"], g["sigma_d"], g["mu"])
    o64 = O.admm_csmri(sd64, g["state"].double(), g["y0"].double(), g["mask"], g["sigma_d"].double(),
                       g["mu"].double())
    assert rel_err(o32, o64)[1] < 1e-5


# ---- CT: parity unpinned -> structural pins ------------------------------------------------

def test_ct_geometry():
    cs, sn, det = O.ct_geometry(256, 60)
    assert det == 363 and len(cs) == 60
    assert abs(float(cs[0]) - 1) < 1e-7 and abs(float(sn[0])) < 1e-7
    assert abs(math.atan2(float(sn[-1]), float(cs[-1])) - 179 / 180 * math.pi) < 1e-6


def test_radon_adjoint_pair():
    n, views = 32, 12
    cs, sn, det = O.ct_geometry(n, views)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 1, n, n, generator=g, dtype=torch.float64)
    y = torch.randn(2, 1, views, det, generator=g, dtype=torch.float64)
    lhs = (O.radon_forward(x, cs, sn, det) * y).sum()
    rhs = (x * O.radon_backward(y, cs, sn, n)).sum()
    assert abs(lhs - rhs) <= 1e-10 * (abs(lhs) + abs(rhs) + 1)


def test_radon_disk_sinogram():
    """Analytic pin: a centred disk of radius r projects to 2*sqrt(r^2 - s^2) at every angle."""
    n, views, r = 64, 8, 20.0
    cs, sn, det = O.ct_geometry(n, views)
    c = (n - 1) / 2
    yy, xx = torch.meshgrid(torch.arange(n) - c, torch.arange(n) - c, indexing="ij")
    img = ((xx ** 2 + yy ** 2) <= r * r).float()[None, None]
    sino = O.radon_forward(img, cs, sn, det)[0, 0]
    s = torch.arange(det) - (det - 1) / 2
    want = 2 * torch.sqrt(torch.clamp(r * r - s ** 2, min=0))
    core = s.abs() < r - 2
    assert ((sino[:, core] - want[core]).abs().max() / (2 * r)) < 0.06
    # mass conservation: every view integrates the image
    assert torch.allclose(sino.sum(1), img.sum().expand(views), rtol=2e-2)


def test_radon_opnorm_seeded():
    cs, sn, det = O.ct_geometry(32, 12)
    a = O.radon_opnorm(32, cs, sn, det, seed=0)
    b = O.radon_opnorm(32, cs, sn, det, seed=0)
    assert a == b and a > 0
    # it is an (under-)estimate of the largest singular value: ||A x|| <= opnorm' ||x|| approx
    x = torch.randn(1, 1, 32, 32, generator=torch.Generator().manual_seed(1))
    assert O.radon_forward(x, cs, sn, det).norm() <= 1.05 * a * x.norm()


def test_ct_loop_runs_and_is_deterministic():
    d = synth.ct_batch(1, 32, 12, 2)
    sd = weights("he")
    a = O.iadmm_ct(sd, d["state"], d["y0"], 12, d["opnorm"], d["sigma_d"], d["mu"], d["tau"])
    b = O.iadmm_ct(sd, d["state"], d["y0"], 12, d["opnorm"], d["sigma_d"], d["mu"], d["tau"])
    assert torch.equal(a, b) and torch.isfinite(a).all() and a.shape == d["state"].shape


# ---- live cross-check against the reference when it is mounted (build container only) --------

@pytest.mark.skipif(not refshim.available(), reason="reference checkout not present (GPU box)")
def test_oracle_matches_live_reference():
    sd = weights("he")
    d = synth.csmri_batch(1, 32, 2, seed=99)
    sol = refshim.reference_solver("csmri", sd)
    with torch.no_grad():
        ref = sol((d["state"], iter((d["y0"], d["mask"]))), (d["sigma_d"], d["mu"]))   # generator aux, as PnPEnv.step
    assert torch.equal(O.admm_csmri(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"]), ref)
