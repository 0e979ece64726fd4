"""IRCNN denoiser (BASELINE configs[0]: csmri ADMM, env_batch=4, 64x64, radial mask, 6 iters, IRCNN).

The reference ships no IRCNN (parity unpinned): the checker is the PyTorch restatement of the published network
(oracle/pnp_oracle.py: ircnn_denoise) on seeded weights.  CPU: the restatement's structure; GPU (-m gpu): the
CUDA path (CUDA-core first/last layers + dilated tcgen05 convolutions) against it.
Tolerances: fp16x3 (split-fp16 emulation of fp32) 1e-4; fp16 (10-bit operand mantissa, = TF32) 2e-3.
"""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import pnp_oracle as O
from oracle import synth


def test_ircnn_oracle_structure():
    sd = synth.ircnn_state_dict(0)
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == O.ircnn_param_shapes()
    assert sum(v.numel() for v in sd.values()) == 186433
    import tfpnp_b200 as T
    assert T.denoiser.ircnn_state_dict_layout() == O.ircnn_param_shapes()
    assert tuple(T.denoiser.IRCNN_DILATIONS) == tuple(O.IRCNN_DILATIONS) == (1, 2, 3, 4, 3, 2, 1)
    x = torch.rand(2, 1, 16, 16); sig = torch.tensor([0.05, 0.2])
    # receptive field of dilations 1,2,3,4,3,2,1 is 33x33: a far-away impulse in the input leaves the pixel unchanged
    y0 = O.ircnn_denoise(sd, x, sig)
    assert y0.shape == x.shape and float(y0.min()) >= 0 and float(y0.max()) <= 1
    zero = {k: torch.zeros_like(v) for k, v in sd.items()}
    assert torch.equal(O.ircnn_denoise(zero, x, sig), x.clamp(0, 1))          # residual form: x - 0
    big = torch.rand(1, 1, 48, 48); big2 = big.clone(); big2[0, 0, 47, 47] += 0.5
    a, b = O.ircnn_denoise(sd, big, sig[:1]), O.ircnn_denoise(sd, big2, sig[:1])
    assert torch.equal(a[0, 0, :30, :30], b[0, 0, :30, :30]) and not torch.equal(a[0, 0, 40:, 40:], b[0, 0, 40:, 40:])


def test_create_denoiser_names():
    import tfpnp_b200 as T

    class Opt:
        denoiser = "ircnn"
    d = T.create_denoiser(Opt(), state_dict=synth.ircnn_state_dict(0))
    assert isinstance(d, T.IRCNNDenoiser2D)
    Opt.denoiser = "dncnn"
    with pytest.raises(NotImplementedError):
        T.create_denoiser(Opt(), state_dict=synth.ircnn_state_dict(0))
    with pytest.raises(KeyError):
        T.IRCNNDenoiser2D(state_dict={})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        d(torch.rand(1, 1, 16, 16), torch.tensor([0.1]))


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16x3", 1e-4), ("fp16", 2e-3)])
@pytest.mark.parametrize("B,H,W", [(2, 32, 32), (3, 64, 48), (4, 64, 64)])
def test_ircnn_denoiser_vs_oracle(dev, prec, tol, B, H, W):
    import tfpnp_b200 as T
    sd = synth.ircnn_state_dict(0)
    g = torch.Generator().manual_seed(B * 100 + H)
    x = torch.rand(B, 1, H, W, generator=g)
    sig = torch.rand(B, generator=g) * (70 / 255)
    ref = O.ircnn_denoise(sd, x, sig)
    den = T.IRCNNDenoiser2D(state_dict=sd, precision=prec)
    out = den(x.to(dev), sig.to(dev))
    torch.cuda.synchronize()
    l2, mx = rel_err(out, ref)
    assert l2 <= tol and mx <= tol, (l2, mx)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,tol", [("fp16x3", 1e-4), ("fp16", 2e-3)])
def test_csmri_config0_ircnn(dev, prec, tol):
    """BASELINE configs[0]: csmri ADMM, env_batch=4, 64x64, radial mask, 6 iters, IRCNN denoiser."""
    import tfpnp_b200 as T
    sd = synth.ircnn_state_dict(0)
    d = synth.csmri_batch(4, 64, 6)
    ref = O.admm_csmri(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"])
    s = T.ADMMSolver_CSMRI(T.IRCNNDenoiser2D(state_dict=sd, precision=prec))
    with torch.no_grad():
        out = s((d["state"].to(dev), (d["y0"].to(dev), d["mask"].to(dev))), (d["sigma_d"].to(dev), d["mu"].to(dev)))
        out2 = s((d["state"].to(dev), (d["y0"].to(dev), d["mask"].to(dev))), (d["sigma_d"].to(dev), d["mu"].to(dev)))
    torch.cuda.synchronize()
    l2, mx = rel_err(out, ref)
    assert l2 <= tol and mx <= tol, (l2, mx)
    assert torch.equal(out, out2)                      # graph replay is deterministic
    p = T.torch_psnr(s.get_output(out), d["gt"].to(dev)).cpu()
    assert torch.allclose(p, O.psnr(O.get_output(ref, True), d["gt"]), atol=5e-2)
