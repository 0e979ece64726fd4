"""GPU parity tests (run with -m gpu on the B200 box): the CUDA path, called through the
reference-facing classes -> ctypes -> C ABI, against the CPU oracle on the same seeded inputs,
against the golden vectors of the unmodified reference, and through size-independent
properties at the BASELINE shapes.

Tolerances (relative to max|ref|, and relative L2), per precision mode of the denoiser convs:
  fp32_simt : 1e-4  (plain fp32 FFMA; observed ~1e-6)
  fp16x3    : 1e-4  (split-fp16 tensor-core emulation of fp32; the mode that meets the
                     north_star 1e-4 contract for ANY weights)
  fp16      : 1e-4 on the SURVEY-prescribed default-init weights; 5e-3 on the variance-preserving
              'he' weights (10-bit operand mantissa = what cuDNN's default TF32 gives the
              reference on a GPU)
"""
import math

import os

import pytest
import torch

from conftest import load_golden, rel_err, weights
from oracle import pnp_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu

PRECS = ["fp32_simt", "fp16x3", "fp16"]


def tol(prec, init):
    if prec == "fp16" and init == "he":
        return 5e-3
    return 1e-4


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


_DEN = {}


def denoiser(prec, init):
    import tfpnp_b200 as T
    key = (prec, init)
    if key not in _DEN:
        _DEN[key] = T.UNetDenoiser2D(state_dict=weights(init), precision=prec)
    return _DEN[key]


def cu(d, dev):
    return {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in d.items()}


def assert_close(got, ref, t, what=""):
    l2, mx = rel_err(got, ref)
    assert math.isfinite(l2) and l2 <= t and mx <= t, f"{what}: relL2 {l2:.3e} relmax {mx:.3e} > {t:.1e}"
    return l2, mx


# ------------------------------------------------------------------------------------------
# denoiser (D1 / D1')
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("init", ["he", "default"])
@pytest.mark.parametrize("prec", PRECS)
def test_denoiser_golden(dev, prec, init):
    g = load_golden(f"denoiser_{init}")
    out = denoiser(prec, init)(g["x"].to(dev), g["sigma"].to(dev))
    assert out.shape == g["out"].shape and out.dtype == torch.float32
    assert_close(out, g["out"], tol(prec, init), f"denoiser {prec}/{init}")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,n", [(3, 64), (2, 128), (5, 16)])
def test_denoiser_vs_oracle_shapes(dev, prec, B, n):
    """ragged batches (partial batch tiles at the 8x8 / 4x4 levels) and the smallest legal size"""
    g = torch.Generator().manual_seed(B * 1000 + n)
    x = torch.rand(B, 1, n, n, generator=g)
    sigma = torch.rand(B, generator=g) * (70 / 255)
    ref = O.denoise(weights("he"), x, sigma)
    out = denoiser(prec, "he")(x.to(dev), sigma.to(dev))
    assert_close(out, ref, tol(prec, "he"), f"denoiser {prec} B={B} n={n}")


def test_denoiser_tc_matches_simt_on_device(dev):
    """tensor-core path vs CUDA-core fp32 path, both on the GPU, larger batch"""
    g = torch.Generator().manual_seed(11)
    x = torch.rand(9, 1, 64, 64, generator=g).to(dev)
    sigma = (torch.rand(9, generator=g) * 0.2).to(dev)
    ref = denoiser("fp32_simt", "he")(x, sigma)
    assert_close(denoiser("fp16x3", "he")(x, sigma), ref, 1e-4, "x3 vs simt")
    assert_close(denoiser("fp16", "he")(x, sigma), ref, 5e-3, "fp16 vs simt")


def test_xform2_transform_warps_are_bit_identical(dev, monkeypatch):
    """TFPNP_XFORM2=1 (row-independent transform warps of the fused up-sampling layers) performs the
    same arithmetic in the same order: the denoiser output must be bit-identical at all three fused levels (32^2, 64^2, 128^2)."""
    g = torch.Generator().manual_seed(5)
    x = torch.rand(3, 1, 128, 128, generator=g).to(dev)
    sigma = (torch.rand(3, generator=g) * 0.2).to(dev)
    den = denoiser("fp16", "he")
    monkeypatch.setenv("TFPNP_XFORM2", "0")
    ref = den(x, sigma).clone()
    monkeypatch.setenv("TFPNP_XFORM2", "1")
    out = den(x, sigma)
    assert torch.equal(out, ref)
    # variant 2: the three lerps in packed fp16 (three roundings instead of one) -- must stay inside the fp16 mode's budget
    monkeypatch.setenv("TFPNP_XFORM2", "2")
    out2 = den(x, sigma)
    oracle = O.denoise(weights("he"), x.cpu(), sigma.cpu())
    e0, e2 = rel_err(ref, oracle)[1], rel_err(out2, oracle)[1]
    print(f"xform2=2: error vs oracle {e2:.2e} (default kernel {e0:.2e})")
    assert e2 <= tol("fp16", "he")


# ------------------------------------------------------------------------------------------
# CS-MRI (S3)
# ------------------------------------------------------------------------------------------

def run_csmri(dev, prec, init, d, **kw):
    import tfpnp_b200 as T
    s = T.ADMMSolver_CSMRI(denoiser(prec, init))
    for k, v in kw.items():
        setattr(s, k, v)
    d = cu(d, dev)
    with torch.no_grad():
        out = s((d["state"], iter((d["y0"], d["mask"]))), (d["sigma_d"], d["mu"]))   # aux as a generator
    return s, out


@pytest.mark.parametrize("prec", PRECS)
def test_csmri_golden_small(dev, prec):
    g = load_golden("csmri_small")
    s, out = run_csmri(dev, prec, "he", g)
    assert out.shape == g["out"].shape
    assert_close(out, g["out"], tol(prec, "he"), f"csmri_small {prec}")
    assert s.last_launch_count > 0


@pytest.mark.parametrize("prec", PRECS)
def test_csmri_golden_cfg1(dev, prec):
    """BASELINE config 1: B=4, 64x64, radial mask, 6 iterations (reference run on CPU)."""
    import tfpnp_b200 as T
    g = load_golden("csmri_cfg1")
    s, out = run_csmri(dev, prec, "default", g)
    assert_close(out, g["out"], tol(prec, "default"), f"csmri_cfg1 {prec}")
    p = T.torch_psnr(s.get_output(out), g["gt"].to(dev))
    assert torch.allclose(p.cpu(), g["psnr"], atol=2e-3)


def test_csmri_mask_selection(dev):
    """Data consistency with mu = 0 replaces exactly the sampled k-space entries by y0 and leaves
    the others alone: pins the roll / sign / permutation folding of y0 and mask."""
    d = synth.csmri_batch(3, 64, 1, seed=5)
    d["mu"] = torch.zeros_like(d["mu"])
    _, out = run_csmri(dev, "fp32_simt", "he", d)
    out = out.cpu()
    x, z, u = torch.split(out, 1, dim=1)
    Z = O.fft2c(z)
    m = d["mask"][..., None].expand_as(Z)
    assert (Z - d["y0"])[m].abs().max() < 2e-5
    # off the mask Z equals fft2c(x + u_old), u_old = 0 initially
    Zin = O.fft2c(x + (u - x + z))
    assert (Z - Zin)[~m].abs().max() < 2e-5
    # and u_new = u_old + x - z with u_old = 0
    assert (u - (x - z)).abs().max() < 1e-6
    assert x[..., 1].abs().max() == 0           # real2complex: exact zero imaginary part


def test_csmri_call_semantics(dev):
    """iter_num < width, non-contiguous parameter columns, shrinking B_left, graph == eager."""
    import tfpnp_b200 as T
    d = synth.csmri_batch(4, 32, 5, seed=3)
    dd = cu(d, dev)
    s = T.ADMMSolver_CSMRI(denoiser("fp16x3", "he"))
    with torch.no_grad():
        full = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"][:, :2].contiguous(), dd["mu"][:, :2].contiguous()))
        part = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"]), iter_num=2)
        assert torch.equal(full, part)
        action = torch.stack([dd["sigma_d"], dd["mu"]], dim=-1)            # [B, it, 2] -> strided views
        strided = s((dd["state"], (dd["y0"], dd["mask"])), (action[..., 0], action[..., 1]), iter_num=2)
        assert torch.equal(full, strided)
        idx = torch.tensor([0, 2], device=dev)                              # env/base.py:162 idx_left gather
        sub = s((dd["state"][idx], (dd["y0"][idx], dd["mask"][idx])), (dd["sigma_d"][idx, :2], dd["mu"][idx, :2]))
        assert torch.equal(sub, full[idx])
        again = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"]), iter_num=2)
        assert torch.equal(again, full)                                     # graph replay after a B change
        s2 = T.ADMMSolver_CSMRI(denoiser("fp16x3", "he"))
        s2.use_graph = False
        eager = s2((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"]), iter_num=2)
        assert torch.equal(eager, full)
        # inputs are not mutated; zero iterations is the identity on (z, u)
        assert torch.equal(dd["state"].cpu(), d["state"])
        zero = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"]), iter_num=0)
        assert torch.equal(zero[:, 1:], dd["state"][:, 1:])
    s.differentiable = False                # gradient requests are refused loudly when the reverse mode is switched off
    with pytest.raises(NotImplementedError):
        with torch.enable_grad():
            s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"].clone().requires_grad_(), dd["mu"]))


def test_forward_validates_shapes_before_marshalling_pointers(dev):
    """Mismatched or broadcast aux tensors raise in the reference (first broadcast); here they must raise on the host
    instead of becoming out-of-bounds device reads (raw pointers cross the C ABI)."""
    import tfpnp_b200 as T
    den = denoiser("fp16", "he")
    d = cu(synth.csmri_batch(3, 32, 2, seed=5), dev)
    s = T.ADMMSolver_CSMRI(den)
    ok = ((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))
    with torch.no_grad():
        s(*ok)
        for bad in (((d["state"], (d["y0"][:2], d["mask"])), ok[1]),                      # fewer y0 rows than images
                    ((d["state"], (d["y0"], d["mask"][:1])), ok[1]),                      # broadcast mask
                    ((d["state"][:, :2], (d["y0"], d["mask"])), ok[1]),                   # two variables, not three
                    ((d["state"], (d["y0"][:, :, :16], d["mask"])), ok[1]),               # wrong spatial size
                    (ok[0], (d["sigma_d"][:2], d["mu"]))):                                # parameter rows != B
            with pytest.raises(ValueError):
                s(*bad)
        h = T.HQSSolver_CSMRI(den)
        st2 = h.reset(dict(x0=d["x0"]))
        h((st2, (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))
        with pytest.raises(ValueError):
            h((d["state"], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))                # 3-variable state into HQS
        with pytest.raises(ValueError):
            h((st2[:, :, :, :16], (d["y0"], d["mask"])), (d["sigma_d"], d["mu"]))         # non-square
        p = cu(synth.pr_batch(2, 32, 2), dev)
        pr = T.IADMMSolver_PR(den)
        pr((p["state"], (p["y0"], p["mask"])), (p["sigma_d"], p["mu"], p["tau"]))
        with pytest.raises(ValueError):
            pr((p["state"], (p["y0"][:, :3], p["mask"])), (p["sigma_d"], p["mu"], p["tau"]))   # 3 magnitudes, 4 masks
        sp = cu(synth.spi_batch(2, 32, 2), dev)
        spi = T.ADMMSolver_SPI(den)
        spi((sp["state"], (sp["x0"], sp["K"])), (sp["sigma_d"], sp["mu"]))
        with pytest.raises(ValueError):
            spi((sp["state"], (sp["x0"][:1], sp["K"])), (sp["sigma_d"], sp["mu"]))
    # re-assigning the denoiser must not reuse a solver handle that captured the old native engine
    s.denoiser = denoiser("fp16x3", "default")
    with torch.no_grad():
        a = s(*ok)
        b = T.ADMMSolver_CSMRI(denoiser("fp16x3", "default"))(*ok)
    assert torch.equal(a, b)


@pytest.mark.parametrize("prec", ["fp16", "fp16x3"])
def test_csmri_full_size_properties(dev, prec):
    """BASELINE config 2 shape (B=48, 128x128): images are independent, so a 48-image call and
    sharded calls must agree BIT FOR BIT (this equality is the multi-GPU test);
    a 12-image slice is also checked against the CPU oracle."""
    import tfpnp_b200 as T
    it = 2
    d = synth.csmri_batch(48, 128, it)
    dd = cu(d, dev)
    s = T.ADMMSolver_CSMRI(denoiser(prec, "default"))
    with torch.no_grad():
        full = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"]))
        parts = []
        for r in range(4):
            lo, hi = T.shard_bounds(48, r, 4)
            sh = T.shard_batch(dd, r, 4)
            parts.append(s((sh["state"], (sh["y0"], sh["mask"])), (sh["sigma_d"], sh["mu"])))
    assert torch.isfinite(full).all()
    assert torch.equal(torch.cat(parts), full)
    sl = slice(0, 12)
    ref = O.admm_csmri(weights("default"), d["state"][sl], d["y0"][sl], d["mask"][sl], d["sigma_d"][sl], d["mu"][sl])
    assert_close(full[sl], ref, 1e-4, f"cfg2 slice {prec}")


# ------------------------------------------------------------------------------------------
# PR (S4), SPI (S6), CT (S5)
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("prec", PRECS)
def test_pr_golden(dev, prec):
    import tfpnp_b200 as T
    g = load_golden("pr_small")
    s = T.IADMMSolver_PR(denoiser(prec, "he"))
    gd = cu(g, dev)
    assert torch.equal(s.reset({"x0": gd["x0"]}).cpu(), g["state"])
    with torch.no_grad():
        out = s((gd["state"], (gd["y0"], gd["mask"])), (gd["sigma_d"], gd["mu"], gd["tau"]))
    assert_close(out, g["out"], tol(prec, "he"), f"pr_small {prec}")


def test_pr_vs_oracle_64(dev):
    import tfpnp_b200 as T
    d = synth.pr_batch(3, 64, 4, seed=21)
    ref = O.iadmm_pr(weights("he"), d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], d["tau"])
    dd = cu(d, dev)
    s = T.IADMMSolver_PR(denoiser("fp16x3", "he"))
    with torch.no_grad():
        out = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"], dd["tau"]))
    assert_close(out, ref, 1e-4, "pr 64")


@pytest.mark.parametrize("prec", PRECS)
def test_spi_golden(dev, prec):
    import tfpnp_b200 as T
    g = load_golden("spi_small")
    s = T.ADMMSolver_SPI(denoiser(prec, "he"))
    gd = cu(g, dev)
    with torch.no_grad():
        out = s((gd["state"], (gd["x0"], gd["K"])), (gd["sigma_d"], gd["mu"]))
    # the 10-step bisection has a 1e-3 resolution: a last-ulp expf difference can flip one branch
    # on isolated pixels, so judge the L2 error and the fraction of deviating pixels
    l2, mx = rel_err(out, g["out"])
    assert l2 <= tol(prec, "he"), (l2, mx)
    if prec != "fp16":
        frac = ((out.cpu() - g["out"]).abs() > 1e-4 * g["out"].abs().max()).float().mean().item()
        assert frac < 1e-3, frac


def test_spi_prox_bit_exact_fraction(dev):
    """one iteration, z slot = spi_inverse(x + u): >= 99.9 % of the pixels bit-identical"""
    import tfpnp_b200 as T
    d = synth.spi_batch(6, 64, 1, seed=8)
    Kv = d["K"][:, 0, 0, 0].reshape(-1, 1, 1, 1) * 10
    x, z, u = torch.split(d["state"], 1, dim=1)
    zref = O.spi_inverse(x + u, d["x0"] * Kv ** 2, Kv, d["mu"][:, 0].reshape(-1, 1, 1, 1))
    dd = cu(d, dev)
    s = T.ADMMSolver_SPI(denoiser("fp32_simt", "he"))
    with torch.no_grad():
        out = s((dd["state"], (dd["x0"], dd["K"])), (dd["sigma_d"], dd["mu"])).cpu()
    zg = out[:, 1:2]
    same = (zg == zref).float().mean().item()
    assert same >= 0.999, same
    assert (zg - zref).abs().max() <= 2.2e-3       # a flipped branch moves by at most one bisection cell
    assert torch.equal(out[:, 2:3], (u + x) - zg)   # u = u + x - z, exact


def test_radon_pair_vs_oracle(dev):
    import tfpnp_b200 as T
    n, views = 64, 24
    cs, sn, det = O.ct_geometry(n, views)
    g = torch.Generator().manual_seed(2)
    img = torch.rand(2, 1, n, n, generator=g)
    sino = torch.randn(2, 1, views, det, generator=g)
    fw = T.radon_forward(img.to(dev), views)
    bw = T.radon_backward(sino.to(dev), n, views)
    assert_close(fw, O.radon_forward(img, cs, sn, det), 1e-5, "radon fwd")
    assert_close(bw, O.radon_backward(sino, cs, sn, n), 1e-5, "radon bwd")
    lhs = (fw.double() * sino.to(dev).double()).sum().item()
    rhs = (img.to(dev).double() * bw.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-5 * (abs(lhs) + abs(rhs))       # <Ax,y> = <x,A^T y>


@pytest.mark.parametrize("n,scale", [(40, 1.0), (64, 1.25), (48, 0.8), (128, 1.0)])
def test_radon_pair_kernel_variants_vs_oracle(dev, n, scale):
    """The three back-projection paths of ct.cu against the oracle's Radon pair: shared-memory windows per 16x16 tile (unit-norm
    tables, N % 16 == 0), the clamped row gather (unit-norm tables, other sizes) and the guarded gather (tables that are NOT unit
    vectors: bins can fall outside the detector); the forward projector's clipped ray walk must add exactly the same terms."""
    from tfpnp_b200 import _lib
    views = 20
    cs, sn, det = O.ct_geometry(n, views)
    cs, sn = (cs * scale).contiguous(), (sn * scale).contiguous()
    g = torch.Generator().manual_seed(5)
    img = torch.rand(2, 1, n, n, generator=g)
    sino = torch.randn(2, 1, views, det, generator=g)
    fw = torch.empty(2, 1, views, det, device=dev)
    bw = torch.empty(2, 1, n, n, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib().tfpnp_radon_forward(img.to(dev).data_ptr(), fw.data_ptr(), 2, n, views, cs.data_ptr(), sn.data_ptr(), st), "fwd")
    _lib.check(_lib.lib().tfpnp_radon_backward(sino.to(dev).data_ptr(), bw.data_ptr(), 2, n, views, cs.data_ptr(), sn.data_ptr(), st), "bwd")
    torch.cuda.synchronize()
    assert_close(fw, O.radon_forward(img, cs, sn, det), 1e-5, f"radon fwd n={n} scale={scale}")
    assert_close(bw, O.radon_backward(sino, cs, sn, n), 1e-5, f"radon bwd n={n} scale={scale}")


@pytest.mark.parametrize("n_masks", [1, 3, 6])
def test_pr_256_mask_counts_vs_oracle(dev, n_masks):
    """pr256_rows_inv splits the masks over four half-warps per row: mask counts that are not a multiple of four."""
    import tfpnp_b200 as T
    d = synth.pr_batch(1, 256, 2, seed=11, n_masks=n_masks)
    ref = O.iadmm_pr(weights("default"), d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], d["tau"])
    dd = cu(d, dev)
    s = T.IADMMSolver_PR(denoiser("fp32_simt", "default"))
    with torch.no_grad():
        out = s((dd["state"], (dd["y0"], dd["mask"])), (dd["sigma_d"], dd["mu"], dd["tau"]))
    assert_close(out, ref, 1e-4, f"pr 256 with {n_masks} masks")


@pytest.mark.parametrize("prec", ["fp32_simt", "fp16x3"])
def test_ct_vs_oracle(dev, prec):
    import tfpnp_b200 as T
    d = synth.ct_batch(2, 64, 24, 3, seed=4)
    ref = O.iadmm_ct(weights("he"), d["state"], d["y0"], 24, d["opnorm"], d["sigma_d"], d["mu"], d["tau"])
    dd = cu(d, dev)
    s = T.IADMMSolver_CT(denoiser(prec, "he"))
    s.opnorm_override = d["opnorm"]
    with torch.no_grad():
        out = s((dd["state"], (dd["y0"], dd["view"])), (dd["sigma_d"], dd["mu"], dd["tau"]))
    assert_close(out, ref, 1e-4, f"ct {prec}")
    # the GPU power method reproduces the seeded CPU operator norm
    s2 = T.IADMMSolver_CT(denoiser(prec, "he"))
    assert abs(s2.radon_generator(64, 24, dev) - d["opnorm"]) <= 1e-4 * d["opnorm"]


def test_psnr_vs_oracle(dev):
    import tfpnp_b200 as T
    g = torch.Generator().manual_seed(1)
    out = torch.rand(7, 1, 128, 128, generator=g) * 1.2 - 0.1
    gt = torch.rand(7, 1, 128, 128, generator=g)
    p = T.torch_psnr(out.to(dev), gt.to(dev))
    assert p.shape == (7, 1)
    assert torch.allclose(p.cpu(), O.psnr(out, gt), rtol=1e-5, atol=1e-4)


def test_native_comm_single_rank(dev):
    """tfpnp_comm_* (NCCL behind the C ABI): a one-rank communicator gathers the PSNR vector onto itself."""
    import tfpnp_b200 as T
    comm = T.NativeComm(T.NativeComm.create_unique_id(), 0, 1, dev)
    p = torch.arange(5, dtype=torch.float32, device=dev).reshape(5, 1) + 0.25
    out = comm.all_gather_psnr(p)
    torch.cuda.synchronize()
    assert torch.equal(out, p)


def test_native_library_is_loaded(dev):
    import tfpnp_b200 as T
    maps = open("/proc/self/maps").read()
    assert "libtfpnp_b200.so" in maps
    assert T.lib().tfpnp_version() == 100


# ------------------------------------------------------------------------------------------
# BASELINE configs 3-5 at their full per-GPU shapes: size-independent properties
# (sharded calls bit-identical to the full call = the multi-GPU equality; finite outputs) plus
# a small slice against the CPU oracle.
# ------------------------------------------------------------------------------------------

def _shard_equal(solver, full, call, n, parts):
    import tfpnp_b200 as T
    outs = []
    for r in range(parts):
        lo, hi = T.shard_bounds(n, r, parts)
        outs.append(call(slice(lo, hi)))
    assert torch.equal(torch.cat(outs), full)


def test_pr_config3_full_shape(dev):
    """pr iADMM, env_batch=36, 256x256, 4 CDP masks (2 iterations here; the loop is linear in iters)."""
    import tfpnp_b200 as T
    d = synth.pr_batch(36, 256, 2)
    dd = cu(d, dev)
    s = T.IADMMSolver_PR(denoiser("fp16", "default"))
    call = lambda sl: s((dd["state"][sl], (dd["y0"][sl], dd["mask"][sl])), (dd["sigma_d"][sl], dd["mu"][sl], dd["tau"][sl]))
    with torch.no_grad():
        full = call(slice(0, 36))
        assert torch.isfinite(full).all()
        _shard_equal(s, full, call, 36, 3)
    sl = slice(0, 2)
    ref = O.iadmm_pr(weights("default"), d["state"][sl], d["y0"][sl], d["mask"][sl], d["sigma_d"][sl], d["mu"][sl], d["tau"][sl])
    assert_close(full[sl], ref, 1e-4, "pr cfg3 slice")


def test_spi_config5_per_gpu_shape(dev):
    """spi ADMM, 48 images per GPU (384 over 8 GPUs), 128x128, K in {4,6,8}."""
    import tfpnp_b200 as T
    d = synth.spi_batch(48, 128, 2)
    dd = cu(d, dev)
    s = T.ADMMSolver_SPI(denoiser("fp16", "default"))
    call = lambda sl: s((dd["state"][sl], (dd["x0"][sl], dd["K"][sl])), (dd["sigma_d"][sl], dd["mu"][sl]))
    with torch.no_grad():
        full = call(slice(0, 48))
        assert torch.isfinite(full).all()
        _shard_equal(s, full, call, 48, 4)
    # The 10-step bisection prox is piecewise constant (1e-3 cells): a denoiser perturbation delta moves a
    # fraction ~delta/1e-3 of the pixels by one cell, so this loop amplifies rounding differences.  The fp32-
    # equivalent mode meets 1e-4; the fp16 mode (10-bit operands, like TF32) is held to 2e-3.
    sl = slice(0, 3)
    ref = O.admm_spi(weights("default"), d["state"][sl], d["x0"][sl], d["K"][sl], d["sigma_d"][sl], d["mu"][sl])
    assert rel_err(full[sl], ref)[0] <= 2e-3
    s3 = T.ADMMSolver_SPI(denoiser("fp16x3", "default"))
    with torch.no_grad():
        out3 = s3((dd["state"][sl], (dd["x0"][sl], dd["K"][sl])), (dd["sigma_d"][sl], dd["mu"][sl]))
    assert rel_err(out3, ref)[0] <= 1e-4


def test_ct_config4_per_gpu_shape(dev):
    """ct iADMM, 8 images per GPU (32 over 4 GPUs), 256x256, 60 views."""
    import tfpnp_b200 as T
    cs, sn, det = O.ct_geometry(256, 60)
    g = torch.Generator().manual_seed(7)
    gt = torch.rand(8, 1, 256, 256, generator=g)
    # measurements through the GPU operators (the CPU restatement of A at 256x256x60 is slow)
    s = T.IADMMSolver_CT(denoiser("fp16", "default"))
    opn = s.radon_generator(256, 60, dev)
    y0 = T.radon_forward(gt.to(dev), 60)
    assert y0.shape == (8, 1, 60, 363)
    x0 = T.radon_backward(y0, 256, 60) / opn ** 2
    state = torch.cat((x0, x0.clone(), torch.zeros_like(x0)), dim=1)
    view = torch.full((8, 1, 256, 256), 60 / 120.0, device=dev)
    par = [(torch.rand(8, 2, generator=g) * sc).to(dev) for sc in (70 / 255, 1.0, 2.0)]
    call = lambda sl: s((state[sl], (y0[sl], view[sl])), tuple(p[sl] for p in par))
    with torch.no_grad():
        full = call(slice(0, 8))
        assert torch.isfinite(full).all() and full.shape == state.shape
        _shard_equal(s, full, call, 8, 2)
    # one image, one iteration against the oracle with the same operator norm
    ref = O.iadmm_ct(weights("default"), state[:1].cpu(), y0[:1].cpu(), 60, opn, par[0][:1, :1].cpu(), par[1][:1, :1].cpu(),
                     par[2][:1, :1].cpu())
    with torch.no_grad():
        one = s((state[:1], (y0[:1], view[:1])), tuple(p[:1, :1] for p in par))
    assert_close(one, ref, 1e-4, "ct cfg4 one image")


# ---- parity at the shapes and iteration counts that bench.py times (BASELINE configs[1..4], all 30 iterations) -------------
# The fp32 floor: this repo's plain-fp32 CUDA-core engine (same update kernels, FFMA convolutions) against the same oracle
# run.  Where the reference loop itself amplifies fp32 rounding -- the unguarded phase division of the PR gradient
# (tasks/pr/solver.py:65-67) and the 1e-3 cells of the SPI bisection (transforms.py:419-437): the CPU oracle in fp32 and in fp64
# differ by 2e-3 there after 30 iterations -- a 1e-4 max-abs bound cannot hold for ANY fp32 implementation, and the tensor-core
# mode is judged against the floor instead.
def _loop_err(dev, task, prec, init, d, n_img, iters, opnorm=None):
    import tfpnp_b200 as T
    sd = weights(init)
    den = denoiser(prec, init)
    sl = slice(0, n_img)
    p = lambda k: d[k][sl, :iters]
    with torch.no_grad():
        if task == "csmri":
            ref = O.admm_csmri(sd, d["state"][sl], d["y0"][sl], d["mask"][sl], p("sigma_d"), p("mu"))
            got = T.ADMMSolver_CSMRI(den)((d["state"][sl].to(dev), (d["y0"][sl].to(dev), d["mask"][sl].to(dev))),
                                          (p("sigma_d").to(dev), p("mu").to(dev)))
        elif task == "pr":
            ref = O.iadmm_pr(sd, d["state"][sl], d["y0"][sl], d["mask"][sl], p("sigma_d"), p("mu"), p("tau"))
            got = T.IADMMSolver_PR(den)((d["state"][sl].to(dev), (d["y0"][sl].to(dev), d["mask"][sl].to(dev))),
                                        (p("sigma_d").to(dev), p("mu").to(dev), p("tau").to(dev)))
        elif task == "spi":
            ref = O.admm_spi(sd, d["state"][sl], d["x0"][sl], d["K"][sl], p("sigma_d"), p("mu"))
            got = T.ADMMSolver_SPI(den)((d["state"][sl].to(dev), (d["x0"][sl].to(dev), d["K"][sl].to(dev))),
                                        (p("sigma_d").to(dev), p("mu").to(dev)))
        else:
            ref = O.iadmm_ct(sd, d["state"][sl], d["y0"][sl], d["views"], opnorm, p("sigma_d"), p("mu"), p("tau"))
            s = T.IADMMSolver_CT(den)
            s.opnorm_override = opnorm
            got = s((d["state"][sl].to(dev), (d["y0"][sl].to(dev), d["view"][sl].to(dev))),
                    (p("sigma_d").to(dev), p("mu").to(dev), p("tau").to(dev)))
    return rel_err(got, ref)


@pytest.mark.parametrize("init", ["default", "he"])
def test_cfg2_csmri_30_iterations(dev, init):
    """BASELINE configs[1]: csmri ADMM, 128x128, all 30 iterations, 4 images of the 48-image batch."""
    d = synth.csmri_batch(48, 128, 30)
    e3 = _loop_err(dev, "csmri", "fp16x3", init, d, 4, 30)
    assert e3[1] <= 1e-4, ("fp16x3", init, e3)                      # the contract mode: 1e-4 (max-abs / max|ref|) for any weights
    e16 = _loop_err(dev, "csmri", "fp16", init, d, 4, 30)
    assert e16[1] <= (1e-4 if init == "default" else 1e-2), ("fp16", init, e16)     # single fp16 products: TF32-class


def test_cfg3_pr_30_iterations(dev):
    """BASELINE configs[2]: pr iADMM, 256x256, 4 CDP masks, all 30 iterations, one image."""
    d = synth.pr_batch(1, 256, 30)
    for init in ("default", "he"):
        floor = _loop_err(dev, "pr", "fp32_simt", init, d, 1, 30)[1]
        e3 = _loop_err(dev, "pr", "fp16x3", init, d, 1, 30)[1]
        assert e3 <= max(1e-4, 4 * floor), (init, e3, floor)


def test_cfg4_ct_3_iterations(dev):
    """BASELINE configs[3] per GPU: ct iADMM, 256x256, 60 views, two images x 3 iterations (own Radon pair: parity unpinned)."""
    d = synth.ct_batch(2, 256, 60, 3)
    for init in ("default", "he"):
        e3 = _loop_err(dev, "ct", "fp16x3", init, d, 2, 3, opnorm=d["opnorm"])
        assert e3[1] <= 1e-4, (init, e3)


def test_cfg5_spi_30_iterations(dev):
    """BASELINE configs[4] per GPU: spi ADMM, 128x128, K in {4,6,8}, all 30 iterations, three images.  The 10-step bisection is
    piecewise constant (1e-3 cells): judged against the fp32 floor and by relative L2."""
    d = synth.spi_batch(48, 128, 30)
    for init in ("default", "he"):
        floor = _loop_err(dev, "spi", "fp32_simt", init, d, 3, 30)
        e3 = _loop_err(dev, "spi", "fp16x3", init, d, 3, 30)
        assert e3[0] <= max(1e-4, 4 * floor[0]) and e3[1] <= max(1e-4, 4 * floor[1]), (init, e3, floor)
        assert e3[0] <= 2e-3, (init, e3)                              # relative L2 stays small even where single cells flip


def test_csmri_fused_update_matches_three_kernel_path(dev):
    """The opt-in one-launch cluster kernel (csmri.cu: csmri_fused) does the same arithmetic in the same order as the
    rows_fwd / cols / rows_inv kernels; only the compiler's FMA contraction inside the FFT butterflies may differ
    between the kernels, so the two paths agree to fp32 rounding (the flag is read once per process -> subprocess)."""
    import subprocess, sys, os
    code = (
        "import sys, torch; sys.path.insert(0, %r);\n"
        "import tfpnp_b200 as T\nfrom oracle import synth\n"
        "dev = torch.device('cuda:0'); sd = synth.unet_state_dict(0, 'default')\n"
        "outs = []\n"
        "for n, B in ((128, 5), (64, 3)):\n"
        "    d = synth.csmri_batch(B, n, 3)\n"
        "    s = T.ADMMSolver_CSMRI(T.UNetDenoiser2D(state_dict=sd, precision='fp16'))\n"
        "    with torch.no_grad():\n"
        "        o = s((d['state'].to(dev), (d['y0'].to(dev), d['mask'].to(dev))), (d['sigma_d'].to(dev), d['mu'].to(dev)))\n"
        "    outs.append(o.cpu())\n"
        "torch.save(outs, sys.argv[1])\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    res = {}
    for flag in ("0", "2", "4"):            # three kernels; one launch on clusters of 2 / of 4 CTAs per image
        path = f"/tmp/csmri_fused_{flag}.pt"
        env = dict(os.environ, TFPNP_CSMRI_FUSED=flag)
        r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        res[flag] = torch.load(path)
    for flag in ("2", "4"):
        for a, b in zip(res["0"], res[flag]):
            assert torch.isfinite(a).all()
            assert_close(b, a, 2e-5, f"fused (cluster of {flag}) vs three-kernel csmri update")
