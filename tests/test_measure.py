"""Stand-alone transforms and GPU measurement synthesis (8f N2).

GPU (-m gpu): tfpnp_fft2 through tfpnp_b200.fft2 / ifft2 / cdp_forward / cdp_backward against the golden vectors
recorded from the unmodified reference (tests/golden/transforms.npz) and against the oracle at every supported size;
the dataset forward models against the oracle's synthetic batches with the noise switched off; an env episode on
GPU-synthesised data.  CPU: the loud failure on CPU tensors and the export of the new symbol.
"""
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import pnp_oracle as O
from oracle import synth


def test_transforms_refuse_cpu():
    import tfpnp_b200 as T
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        T.fft2(torch.zeros(1, 1, 32, 32, 2))
    assert hasattr(T.lib(), "tfpnp_fft2")


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
def test_transforms_golden(dev):
    import tfpnp_b200 as T
    g = load_golden("transforms")
    x, m, gg = g["x"].to(dev), g["mask"].to(dev), g["g"].to(dev)
    assert rel_err(T.fft2(x), g["fft2"])[1] <= 2e-6
    assert rel_err(T.ifft2(x), g["ifft2"])[1] <= 2e-6
    assert rel_err(T.cdp_forward(x, m), g["cdp_fwd"])[1] <= 2e-6
    assert rel_err(T.cdp_backward(gg, m), g["cdp_bwd"])[1] <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n", [32, 64, 128, 256])
def test_fft2_all_sizes_and_round_trip(dev, n):
    import tfpnp_b200 as T
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(3, 2, n, n, 2, generator=gen)
    X = T.fft2(x.to(dev))
    assert rel_err(X, O.fft2c(x))[1] <= 2e-6
    assert rel_err(T.ifft2(x.to(dev)), O.ifft2c(x))[1] <= 2e-6
    assert rel_err(T.ifft2(X), x)[1] <= 2e-6                       # round trip
    # Parseval (ortho): energy preserved
    assert abs(float((X ** 2).sum()) / float((x ** 2).sum()) - 1) <= 1e-5
    m = synth.pr_batch(3, n, 1)["mask"]
    xr = x[:, :1]
    assert rel_err(T.cdp_forward(xr.to(dev), m.to(dev)), O.cdp_forward(xr, m))[1] <= 2e-6


@pytest.mark.gpu
def test_csmri_and_pr_measurements_match_dataset_formulas(dev):
    import tfpnp_b200 as T
    d = synth.csmri_batch(4, 64, 1, sigma_n=0.0)                    # noise-free reference formulas
    out = T.csmri_measure(d["gt"].to(dev), d["mask"].to(dev), sigma_n=0.0)
    assert rel_err(out["y0"], d["y0"])[1] <= 2e-6 and rel_err(out["x0"], d["x0"])[1] <= 2e-6
    assert torch.equal(out["mask"].cpu(), d["mask"].bool())
    assert float((out["y0"] * (~out["mask"])[..., None]).abs().max()) == 0.0      # exactly zero off the mask
    assert rel_err(out["output"], O.complex2real(d["x0"]))[1] <= 2e-6
    noisy = T.csmri_measure(d["gt"].to(dev), d["mask"].to(dev), sigma_n=15 / 255, generator=torch.Generator(dev).manual_seed(1))
    resid = (noisy["y0"] - out["y0"])[out["mask"].expand(-1, -1, -1, -1)]
    assert abs(float(resid.std()) / (15 / 255) - 1) < 0.05
    p = synth.pr_batch(3, 64, 1, alpha=0.0)
    pm = T.pr_measure(p["gt"].to(dev), p["mask"].to(dev), alpha=0.0)
    assert rel_err(pm["y0"], p["y0"])[1] <= 2e-6 and torch.equal(pm["x0"].cpu(), p["x0"])


@pytest.mark.gpu
def test_episode_on_gpu_synthesised_data(dev):
    """N2 -> N1 -> hot path: synthesise CS-MRI measurements on the GPU, run a 2-step episode, PSNR must improve on the
    zero-filled reconstruction for a denoiser that is close to the identity (zero last layer)."""
    import tfpnp_b200 as T
    sd = synth.unet_state_dict(0, "default")
    sd = {k: (torch.zeros_like(v) if k.startswith("outc") else v) for k, v in sd.items()}   # UNet(x) = x
    gt = torch.rand(4, 1, 64, 64, generator=torch.Generator().manual_seed(3)).to(dev)
    mask = torch.stack([synth.radial_mask(64, 20)] * 4)[:, None].to(dev)
    data = T.csmri_measure(gt, mask, sigma_n=5 / 255, generator=torch.Generator(dev).manual_seed(2))
    env = T.CSMRIEnv(None, T.ADMMSolver_CSMRI(T.UNetDenoiser2D(state_dict=sd, precision="fp16x3")), 2).to(dev)
    ob = env.reset(data=data)
    assert env.get_policy_ob(ob).shape == (4, 9, 64, 64)
    psnr0 = env.last_metric.clone()
    total = torch.zeros_like(psnr0)
    for _ in range(2):
        a = {"sigma_d": torch.full((4, 3), 5 / 255, device=dev), "mu": torch.full((4, 3), 0.5, device=dev),
             "idx_stop": torch.zeros(4, dtype=torch.long, device=dev)}
        ob, ob_m, reward, all_done, info = env.step(a)
        total += reward
    assert all_done and torch.isfinite(total).all()
    assert torch.allclose(env.last_metric, psnr0 + total, atol=1e-3)


def test_radial_mask_and_seeded_weights_cpu():
    """Host-side helpers of the benchmark inputs run anywhere (no kernels involved)."""
    import tfpnp_b200 as T
    from oracle import pnp_oracle as O
    for n, lines, lo, hi in ((128, 42, 0.30, 0.45), (128, 21, 0.15, 0.25), (64, 10, 0.10, 0.22)):
        m = T.radial_mask(n, lines)
        assert m.dtype == torch.bool and tuple(m.shape) == (n, n)
        assert lo <= float(m.float().mean()) <= hi
        assert bool(m[n // 2 - 1:n // 2 + 1, n // 2 - 1:n // 2 + 1].any())      # lines pass through the centre
    sd = T.random_unet_state_dict(3)
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == O.unet_param_shapes()
    sd2 = T.random_unet_state_dict(3)
    assert all(torch.equal(sd[k], sd2[k]) for k in sd)                          # seeded
    w = sd["inc.conv.conv-1.conv2d.weight"]
    assert float(w.abs().max()) <= 1 / (32 * 9) ** 0.5 + 1e-7                   # U(+-1/sqrt(fan_in)), nn.Conv2d's default
