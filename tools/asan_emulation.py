"""Driver of tools/asan_emulation.sh: every emulated CUDA sequence of the reverse mode (tests/grad_elem_host.cpp) once, including
the smallest legal image (16x16: the deepest level is 1x1) and a ragged batch, under AddressSanitizer + UBSan."""
import ctypes as C, os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden, weights
from oracle import grad_oracle as G, pnp_oracle as O
from tfpnp_b200.denoiser import flatten_state_dict
emu=C.CDLL(os.path.join(ROOT, 'gpurun_out', 'asan', 'emu_asan.so'))
p=lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
sd=weights("he"); flat=flatten_state_dict(sd)
g=load_golden("grad_csmri_small")
x,s,go=g["den_x"].contiguous(),g["den_sigma"].contiguous(),g["den_gout"].contiguous()
for mode in (0,1,2,3):
    gx=torch.zeros_like(x); gs=torch.zeros(2)
    print("vjp mode",mode, emu.emu_unet_vjp_tc(p(flat),p(x),p(s),p(go),p(gx),p(gs),2,32,32,mode) if mode else emu.emu_unet_vjp(p(flat),p(x),p(s),p(go),p(gx),p(gs),2,32,32))
# smallest legal size 16x16 (level 4 is 1x1) and a ragged batch
x16=torch.rand(3,1,16,16); s16=torch.rand(3)*0.2; go16=torch.randn(3,1,16,16); gx=torch.zeros_like(x16); gs=torch.zeros(3)
print("vjp 16x16 B=3", emu.emu_unet_vjp(p(flat),p(x16),p(s16),p(go16),p(gx),p(gs),3,16,16), emu.emu_unet_vjp_tc(p(flat),p(x16),p(s16),p(go16),p(gx),p(gs),3,16,16,2))
rx,rs=G.denoise_vjp_autograd(sd,x16,s16,go16); print("  16x16 err", ((gx-rx).norm()/rx.norm()).item())
st=torch.stack(G.admm_csmri_trajectory(sd,g["state"],g["y0"],g["mask"],g["sigma_d"],g["mu"])).contiguous()
m8=g["mask"].to(torch.uint8).contiguous(); B,it=g["sigma_d"].shape
a,b,c=torch.zeros(B,it),torch.zeros(B,it),torch.zeros_like(g["gout"])
print("admm", emu.emu_admm_backward(p(flat),p(st),p(g["y0"]),p(m8),p(g["sigma_d"]),p(g["mu"]),B,32,it,p(g["gout"]),p(a),p(b),p(c)))
gs_=load_golden("grad_spi_small"); st=torch.stack(G.admm_spi_trajectory(sd,gs_["state"],gs_["x0"],gs_["K"],gs_["sigma_d"],gs_["mu"])).contiguous()
Kv=gs_["K"][:,0,0,0].contiguous(); B,it=gs_["sigma_d"].shape
a,b,c=torch.zeros(B,it),torch.zeros(B,it),torch.zeros_like(gs_["gout"])
print("spi", emu.emu_spi_backward(p(flat),p(st),p(gs_["x0"]),p(Kv),C.c_int64(1),p(gs_["sigma_d"]),p(gs_["mu"]),B,32,32,it,p(gs_["gout"]),p(a),p(b),p(c)))
gp=load_golden("grad_pr_small"); st=torch.stack(G.iadmm_pr_trajectory(sd,gp["state"],gp["y0"],gp["mask"],gp["sigma_d"],gp["mu"],gp["tau"])).contiguous()
B,it=gp["sigma_d"].shape; o=[torch.zeros(B,it) for _ in range(3)]+[torch.zeros_like(gp["gout"])]
print("pr", emu.emu_pr_backward(p(flat),p(st),p(gp["y0"]),p(gp["mask"]),4,p(gp["sigma_d"]),p(gp["mu"]),p(gp["tau"]),B,32,it,p(gp["gout"]),*[p(t) for t in o]))
gv=load_golden("grad_csmri_variants")
for name,algo,fn,keys in (("hqs",1,O.hqs_csmri,("sigma_d","mu")),("pg",2,O.pg_csmri,("sigma_d","tau")),("apg",3,O.apg_csmri,("sigma_d","tau","beta")),("redadmm",4,O.redadmm_csmri,("sigma_d","mu","lamda"))):
    ps=[gv[k] for k in keys]; states=[gv[name+"_state0"]]
    with torch.no_grad():
        for i in range(3): states.append(fn(sd,states[-1],gv["y0"],gv["mask"],*[q[:,i:i+1] for q in ps]))
    S=torch.stack(states).contiguous(); m8=gv["mask"].to(torch.uint8).contiguous()
    o=[torch.zeros(2,3) for _ in range(3)]; gst=torch.zeros_like(states[0]); pp=[q.contiguous() for q in ps]+[None]*(3-len(ps)); cot=gv[name+"_gout"].contiguous()
    print(name, emu.emu_variant_backward(algo,p(flat),p(S),p(gv["y0"]),p(m8),p(pp[0]),p(pp[1]),p(pp[2]),2,32,3,p(cot),p(o[0]),p(o[1]),p(o[2]),p(gst)))
print("asan run complete")
