#!/usr/bin/env python
"""Role timeline of CTA 0 for one UNet layer inside a real denoiser call (TFPNP_TRACE_LAYER)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
names = ["P: A slot free", "M: tile start", "M: tmem free", "M: A landed", "M: tile issued", "X: window landed / stage written",
         "E: accum ready", "E: tile done"]
if len(sys.argv) > 2 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import numpy as np, torch, tfpnp_b200 as T
    from oracle import synth
    layer = int(sys.argv[2])
    dev = torch.device("cuda:0")
    den = T.UNetDenoiser2D(state_dict=synth.unet_state_dict(0, "default"), precision="fp16")
    x = torch.rand(48, 1, 128, 128, device=dev); sig = torch.full((48,), 0.1, device=dev)
    for _ in range(3):
        den(x, sig)
    torch.cuda.synchronize()
    tr = np.fromfile("/tmp/trace.bin", dtype=np.uint64).reshape(8, 1024).astype(np.int64)
    t0 = tr[tr > 0].min()
    extra = tr[0, 1000:1016].copy(); tr[0, 1000:] = 0
    print(f"=== layer {layer}: entry, pre-sync, post-sync(pdl), final-sync, dealloc:", [int(v - t0) if v else None for v in extra[:5]],
          " role loops done (warps 0..5):", [int(v - t0) if v else None for v in extra[10:16]])
    for r in range(8):
        row = tr[r][tr[r] > 0] - t0
        print(f"  {names[r]:34s} n={len(row):3d}: {row[:26].tolist()} ... {row[-2:].tolist()}")
else:
    for layer in sys.argv[1:]:
        env = dict(os.environ, TFPNP_TRACE_LAYER=layer, TFPNP_TRACE_FILE="/tmp/trace.bin")
        out = subprocess.run([sys.executable, __file__, "child", layer], env=env, capture_output=True, text=True)
        print(out.stdout[-6000:] if out.stdout else out.stderr[-2000:])
