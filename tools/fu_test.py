import sys, torch
sys.path.insert(0, '/root/repo')
import tfpnp_b200 as T
from oracle import synth, pnp_oracle as O
sd = synth.unet_state_dict(0, 'he')
den = T.UNetDenoiser2D(state_dict=sd, precision='fp16')
g = torch.Generator().manual_seed(1)
for B, n in [(1, 32), (2, 64)]:
    x = torch.rand(B, 1, n, n, generator=g); s = torch.rand(B, generator=g) * 0.2
    out = den(x.cuda(), s.cuda()); torch.cuda.synchronize()
    ref = O.denoise(sd, x, s)
    print(B, n, 'relmax', ((out.cpu() - ref).abs().max() / ref.abs().max()).item())
