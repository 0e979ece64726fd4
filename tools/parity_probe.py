import sys, torch
sys.path.insert(0, "/root/repo")
import tfpnp_b200 as T
from oracle import pnp_oracle as O, synth
dev = torch.device("cuda:0")
def rel(a, b): return ((a - b).abs().max() / b.abs().max()).item()
d = synth.pr_batch(1, 256, 30)
for init in ("default", "he"):
    sd = synth.unet_state_dict(0, init)
    for prec in ("fp32_simt", "fp16x3"):
        s = T.IADMMSolver_PR(T.UNetDenoiser2D(state_dict=sd, precision=prec))
        row = []
        for it in (1, 2, 5, 10, 20, 30):
            ref = O.iadmm_pr(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], d["tau"], iter_num=it)
            with torch.no_grad():
                got = s((d["state"].to(dev), (d["y0"].to(dev), d["mask"].to(dev))), (d["sigma_d"].to(dev), d["mu"].to(dev), d["tau"].to(dev)), iter_num=it).cpu()
            row.append(f"{it}:{rel(got, ref):.1e}")
        print("pr", init, prec, " ".join(row), flush=True)
d = synth.csmri_batch(4, 128, 30)
for init in ("he",):
    sd = synth.unet_state_dict(0, init)
    for prec in ("fp32_simt", "fp16x3"):
        s = T.ADMMSolver_CSMRI(T.UNetDenoiser2D(state_dict=sd, precision=prec))
        row = []
        for it in (1, 5, 10, 20, 30):
            ref = O.admm_csmri(sd, d["state"], d["y0"], d["mask"], d["sigma_d"], d["mu"], iter_num=it)
            with torch.no_grad():
                got = s((d["state"].to(dev), (d["y0"].to(dev), d["mask"].to(dev))), (d["sigma_d"].to(dev), d["mu"].to(dev)), iter_num=it).cpu()
            row.append(f"{it}:{rel(got, ref):.1e}")
        print("csmri", init, prec, " ".join(row), flush=True)
# single denoiser call error
x = torch.rand(4, 1, 128, 128); sg = torch.rand(4) * 0.2
for init in ("default", "he"):
    sd = synth.unet_state_dict(0, init)
    ref = O.denoise(sd, x, sg)
    ref64 = O.denoise({k: v.double() for k, v in sd.items()}, x.double(), sg.double())
    print("denoise", init, "oracle32 vs 64", rel(ref.double(), ref64))
    for prec in ("fp32_simt", "fp16x3", "fp16"):
        den = T.UNetDenoiser2D(state_dict=sd, precision=prec)
        with torch.no_grad():
            got = den(x.to(dev), sg.to(dev)).cpu()
        print("denoise", init, prec, "vs oracle32", rel(got, ref), "vs oracle64", rel(got.double(), ref64))
