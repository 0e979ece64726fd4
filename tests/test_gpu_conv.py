"""Kernel-level parity of the tcgen05 implicit-GEMM convolution (one UNet ConvLayer,
tfpnp/pnp/denoiser/models/unet.py:8-22) against torch's fp32 CPU convolution on the same
fp16-representable operands.  Covers both swizzle widths (32- and 64-channel chunks), every
N tile (32/64/128, multi-N-tile), the two-source concat K loop, zero padding at all borders,
ragged batch tiles and the small-resolution tile geometries."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # C0, C1, Cout, H,  W,  B
    (32, 0, 32, 16, 16, 1),
    (32, 0, 32, 32, 24, 2),
    (64, 0, 64, 16, 16, 2),
    (32, 0, 64, 16, 8, 1),
    (32, 64, 32, 16, 16, 1),      # up4.conv-0: cat[skip 32, up 64], 32-channel chunks
    (64, 128, 64, 16, 16, 1),     # up3.conv-0
    (128, 0, 128, 16, 16, 1),
    (128, 0, 256, 16, 8, 2),      # two N tiles
    (256, 512, 256, 16, 16, 1),   # up1.conv-0, K = 6912
    (256, 0, 512, 8, 8, 3),       # down4: 2 images per tile, ragged batch
    (64, 0, 64, 4, 4, 3),         # 4x4 level of a 64x64 image: 8 images per tile
    (512, 0, 512, 2, 2, 5),       # 2x2 level of a 32x32 image: 32 images per tile
    (512, 0, 512, 1, 1, 5),       # 1x1 level of a 16x16 image
    (96, 0, 32, 32, 32, 2),       # three 32-channel chunks from one source
    (32, 0, 32, 128, 128, 4),     # 512 tiles > resident CTAs: persistent multi-tile loop, TMEM double buffer
    (64, 0, 64, 64, 64, 8),       # resident 72 KB weights, 1 CTA/SM
    (128, 0, 128, 32, 32, 6),     # streamed weights, multi-tile
    (64, 128, 64, 64, 64, 3),     # streamed weights, concat, 3 chunks
    (512, 0, 512, 8, 8, 48),      # 8x8 level at the bench shape: two-image M-tiles, 4 N tiles, weight multicast
    (256, 512, 256, 8, 8, 5),     # 8x8 level, concat of two sources, ragged (odd) batch
    (512, 0, 512, 8, 8, 1),       # a single 8x8 image: falls back to the one-tile-per-CTA kernel
    (32, 0, 64, 64, 64, 3),       # down1.conv-0: 32-channel chunk, N = 64, resident weights, 2 CTAs / SM
    (64, 0, 64, 64, 64, 48),      # the bench shape of the 64-channel level (768 tiles on 148 persistent CTAs)
]


@pytest.mark.parametrize("C0,C1,Cout,H,W,B", CASES)
def test_conv3x3_tc_vs_torch(C0, C1, Cout, H, W, B):
    import tfpnp_b200 as T
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(C0 * 7 + C1 * 3 + Cout + H + W + B)
    x0 = torch.randn(B, H, W, C0, generator=g).half()
    x1 = torch.randn(B, H, W, C1, generator=g).half() if C1 else None
    cin = C0 + C1
    w = (torch.randn(Cout, cin, 3, 3, generator=g) * (2.0 / (9 * cin)) ** 0.5).half().float()
    b = torch.randn(Cout, generator=g) * 0.1
    xin = x0.float() if x1 is None else torch.cat([x0.float(), x1.float()], dim=-1)
    ref = F.leaky_relu(F.conv2d(xin.permute(0, 3, 1, 2), w, b, padding=1), 0.2).permute(0, 2, 3, 1)
    out = T.conv3x3_lrelu_nhwc(x0.to(dev), w, b, None if x1 is None else x1.to(dev))
    torch.cuda.synchronize()
    assert out.shape == ref.shape
    l2, mx = rel_err(out.float(), ref)
    assert l2 < 6e-4 and mx < 2e-3, (l2, mx)      # fp16 output rounding: 2^-11 relative


PAIR_CASES = [(128, 0, 128, 32, 32, 6), (256, 0, 256, 16, 16, 5), (256, 512, 256, 16, 16, 2), (64, 0, 128, 32, 16, 3)]


@pytest.mark.parametrize("C0,C1,Cout,H,W,B", PAIR_CASES)
def test_conv3x3_pair_kernel_vs_torch(C0, C1, Cout, H, W, B, monkeypatch):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2) and, with TFPNP_CONV_PAIR=0, the single-CTA kernel it
    replaces on these shapes: both against torch."""
    monkeypatch.setenv("TFPNP_CONV_PAIR", "1")
    test_conv3x3_tc_vs_torch(C0, C1, Cout, H, W, B)
    monkeypatch.setenv("TFPNP_CONV_PAIR", "0")
    test_conv3x3_tc_vs_torch(C0, C1, Cout, H, W, B)
