#!/usr/bin/env python
"""Bottleneck hunting for the persistent conv kernel: time one layer shape with parts of the
pipeline knocked out (TFPNP_DBG bits: 1 no stores, 2 no MMA, 4 no activation TMA, 8 no weight TMA)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(32, 0, 32, 128, 128, 48), (64, 0, 64, 64, 64, 48), (128, 0, 128, 32, 32, 48), (256, 0, 256, 16, 16, 48),
          (512, 0, 512, 8, 8, 48), (64, 128, 64, 64, 64, 48), (32, 64, 32, 128, 128, 48), (256, 512, 256, 16, 16, 48)]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch, tfpnp_b200 as T
    dev = torch.device("cuda:0")
    res = {}
    for (C0, C1, Cout, H, W, B) in SHAPES:
        x0 = torch.randn(B, H, W, C0, device=dev).half()
        x1 = torch.randn(B, H, W, C1, device=dev).half() if C1 else None
        w = torch.randn(Cout, C0 + C1, 3, 3) * 0.05
        b = torch.zeros(Cout)
        for _ in range(3):
            T.conv3x3_lrelu_nhwc(x0, w, b, x1)
        torch.cuda.synchronize()
        wt = w.permute(2, 3, 0, 1).reshape(9, Cout, C0 + C1).half().contiguous().to(dev)
        bb = b.to(dev); out = torch.empty(B, H, W, Cout, device=dev, dtype=torch.float16)
        from tfpnp_b200 import _lib
        st = torch.cuda.current_stream().cuda_stream
        # GPU-side time: 20 launches captured in a CUDA graph (host-side tensor-map encoding and launch
        # latency would otherwise hide any kernel shorter than ~25 us), median of 5 replays
        cs = torch.cuda.Stream()
        def launch(stream):
            _lib.check(_lib.lib().tfpnp_conv3x3_nhwc(x0.data_ptr(), C0, x1.data_ptr() if x1 is not None else None, C1,
                                                     wt.data_ptr(), bb.data_ptr(), out.data_ptr(), B, H, W, Cout, stream), "conv")
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(cs):
            launch(cs.cuda_stream); torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=cs):
                for _ in range(20):
                    launch(cs.cuda_stream)
        ts = []
        for _ in range(5):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); g.replay(); e.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(e) / 20)
        t = sorted(ts)[2] * 1e3
        gf = 2 * 9 * (C0 + C1) * Cout * H * W * B / 1e9
        res[f"{C0}+{C1}->{Cout}@{H}"] = (round(t, 1), round(gf / t * 1e3, 1))
    print(json.dumps(res))
else:
    for dbg in ([0, 1, 2, 3, 4, 8, 15] if os.environ.get("KNOCK_ALL") else [0, 32, 15, 128, 64]):
        env = dict(os.environ, TFPNP_DBG=str(dbg))
        out = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        print(f"dbg={dbg:2d} (us, TFLOP/s): {line}")
