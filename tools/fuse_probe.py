import sys, os, torch
sys.path.insert(0, "/root/repo")
import tfpnp_b200 as T
from oracle import pnp_oracle as O, synth
dev = torch.device("cuda:0")
sd = synth.unet_state_dict(0, "he")
den = T.UNetDenoiser2D(state_dict=sd, precision="fp16x3")
g = torch.Generator().manual_seed(1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
x = torch.rand(48, 1, N, N, generator=g).to(dev); sg = (torch.rand(48, generator=g) * 0.2).to(dev)
with torch.no_grad():
    full = den(x, sg)
    for n in (12, 5, 48):
        part = den(x[:n].contiguous(), sg[:n].contiguous())
        d = (part - full[:n]).abs()
        print("B", n, "max diff", d.max().item(), "n diff px", int((d > 0).sum()))
        if d.max() > 0:
            idx = (d > 0).nonzero()
            print("  images", sorted(set(idx[:, 0].tolist()))[:10], "rows", sorted(set(idx[:, 2].tolist()))[:40], "cols", sorted(set(idx[:, 3].tolist()))[:40])
    again = den(x, sg)
    print("repeat full: max diff", (again - full).abs().max().item())
    ref = O.denoise(sd, x[:4].cpu(), sg[:4].cpu())
    print("vs oracle", ((full[:4].cpu() - ref).abs().max() / ref.abs().max()).item())
