#!/usr/bin/env python
"""Per-role timeline of CTA 0 of the persistent conv kernel (globaltimer samples)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, tfpnp_b200 as T
dev = torch.device("cuda:0")
names = ["P: A slot free", "M: tile start", "M: tmem free", "M: A landed", "M: tile issued", "E: tile wait", "E: accum ready", "E: tile done"]
for (C0, C1, Cout, H, W, B) in [(64, 0, 64, 64, 64, 48), (32, 0, 32, 128, 128, 48), (128, 0, 128, 32, 32, 48), (256, 0, 256, 16, 16, 48)]:
    x0 = torch.randn(B, H, W, C0, device=dev).half()
    w = torch.randn(Cout, C0 + C1, 3, 3) * 0.05
    b = torch.zeros(Cout)
    for _ in range(2):
        T.conv3x3_lrelu_nhwc(x0, w, b)
    torch.cuda.synchronize()
    os.environ["TFPNP_TRACE_FILE"] = "/tmp/trace.bin"
    T.conv3x3_lrelu_nhwc(x0, w, b)
    torch.cuda.synchronize()
    del os.environ["TFPNP_TRACE_FILE"]
    tr = np.fromfile("/tmp/trace.bin", dtype=np.uint64).reshape(8, 1024).astype(np.int64)
    t0 = tr[tr > 0].min()
    extra = tr[0, 1000:1016].copy(); tr[0, 1000:] = 0
    print("  kernel entry, prologue pre-sync, post-sync, post-final-sync, post-dealloc:", [int(x - t0) if x else None for x in extra[:5]],
          " role loops done (warps 0..5):", [int(x - t0) if x else None for x in extra[10:16]])
    print(f"=== {C0}->{Cout} @{H}x{W} B={B}  (ns since first sample; first 8 tiles and last 2)")
    fine = tr[5][:60] - t0
    print("  resident MMA issue: time before each tap's MMAs (9 taps) + after the block, tiles 1..4 (ia=1..4):")
    for k in range(1, 5):
        print("     ", fine[10 * k: 10 * k + 10].tolist())
    tr[5] = 0
    for r in range(8):
        row = tr[r][tr[r] > 0] - t0
        print(f"  {names[r]:18s} n={len(row):3d}: {row[:8].tolist()} ... {row[-2:].tolist()}")
