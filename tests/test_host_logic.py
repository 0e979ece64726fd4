"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header
declares, the reference-facing classes keep the reference's interface and error behaviour,
and the multi-process (gloo, world_size 2) sharding / PSNR gather logic."""
import os
import re
import subprocess
import sys
import types

import pytest
import torch

from conftest import ROOT, weights

import tfpnp_b200 as T
from tfpnp_b200 import _lib


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "tfpnp_b200.h")).read()
    declared = set(re.findall(r"\b(tfpnp_[a-z_0-9]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    l = _lib.lib()                       # dlopen + symbol resolution; no compute
    for name in declared:
        assert hasattr(l, name)
    assert l.tfpnp_version() == 100


def test_ctypes_signatures_match_the_header_types():
    """Every parameter of every declaration in include/tfpnp_b200.h against the ctypes argtypes in tfpnp_b200/_lib.py: pointer
    -> c_void_p / POINTER(...), int -> c_int, int64_t -> c_int64, size_t -> c_size_t, float -> c_float (a c_int bound to an
    int64_t stride would pass garbage in the upper half on this ABI only by luck)."""
    import ctypes as C
    header = open(os.path.join(ROOT, "include", "tfpnp_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    decls = re.findall(r"\b([a-z_0-9 ]+?\**)\s*\b(tfpnp_[a-z_0-9]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)
    assert len(decls) == len(_lib.SIGNATURES), (len(decls), len(_lib.SIGNATURES))

    def kind(ctype_decl):
        t = " ".join(ctype_decl.replace("const", " ").split())
        if "*" in t:
            return "ptr"
        base = t.split()[0] if t.split() else "void"
        return {"int": "int", "int64_t": "i64", "size_t": "size", "float": "float", "void": "void"}.get(base, base)

    def ckind(a):
        if a is C.c_void_p or a is C.c_char_p or (isinstance(a, type) and issubclass(a, C._Pointer)):
            return "ptr"
        return {C.c_int: "int", C.c_int64: "i64", C.c_size_t: "size", C.c_float: "float"}[a]

    for ret, name, params in decls:
        res, args = _lib.SIGNATURES[name]
        plist = [q.strip() for q in params.split(",")] if params.strip() not in ("", "void") else []
        kinds = []
        for q in plist:
            q = re.sub(r"\b[A-Za-z_][A-Za-z_0-9]*$", "", q).strip() if not q.endswith("*") else q      # drop the parameter name
            kinds.append(kind(q))
        assert kinds == [ckind(a) for a in args], (name, kinds, [ckind(a) for a in args])
        assert kind(ret) == ("ptr" if res is C.c_char_p else ckind(res)), (name, ret)


def test_library_is_sm100a_tcgen05():
    """The shipped cubin is sm_100a and contains the Blackwell tensor/TMA instructions."""
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in out.stdout, mnemonic


def test_flatten_state_dict_validates():
    from tfpnp_b200.denoiser import flatten_state_dict, unet_state_dict_layout
    sd = weights("he")
    flat = flatten_state_dict(sd)
    assert flat.numel() == 11773857 and flat.dtype == torch.float32
    assert [k for k, _ in unet_state_dict_layout()] == list(sd.keys())
    bad = dict(sd)
    bad.pop("outc.conv.bias")
    with pytest.raises(KeyError):
        flatten_state_dict(bad)
    bad = dict(sd)
    bad["inc.conv.conv-0.conv2d.weight"] = torch.zeros(32, 3, 3, 3)
    with pytest.raises(ValueError):
        flatten_state_dict(bad)


def test_interface_mirrors_reference():
    den = T.UNetDenoiser2D(state_dict=weights("he"))
    with pytest.raises(ValueError):
        T.UNetDenoiser2D()               # denoiser/base.py:10-13
    s = T.ADMMSolver_CSMRI(den)
    assert s.num_var == 3 and isinstance(s, torch.nn.Module) and s.denoiser is den
    x0 = torch.randn(2, 1, 8, 8, 2)
    st = s.reset({"x0": x0})
    assert st.shape == (2, 3, 8, 8, 2) and torch.equal(st[:, 0], x0[:, 0]) and torch.equal(st[:, 1], x0[:, 0])
    assert torch.count_nonzero(st[:, 2]) == 0
    assert torch.equal(s.get_output(st), x0[..., 0])
    act = {"sigma_d": 1, "mu": 2, "tau": 3, "idx_stop": 4}
    assert s.filter_hyperparameter(act) == (1, 2)
    assert T.IADMMSolver_PR(den).filter_hyperparameter(act) == (1, 2, 3)
    assert T.IADMMSolver_CT(den).filter_hyperparameter(act) == (1, 2, 3)
    state = {"y0": "y", "mask": "m", "view": "v", "x0": "x", "K": "k"}
    assert s.filter_aux_inputs(state) == ("y", "m")
    assert T.IADMMSolver_PR(den).filter_aux_inputs(state) == ("y", "m")
    assert T.IADMMSolver_CT(den).filter_aux_inputs(state) == ("y", "v")
    assert T.ADMMSolver_SPI(den).filter_aux_inputs(state) == ("x", "k")
    pr = T.IADMMSolver_PR(den).reset({"x0": torch.ones(1, 1, 4, 4)})
    assert pr.shape == (1, 3, 4, 4, 2) and torch.equal(pr[:, 0, ..., 0], torch.ones(1, 4, 4))
    spi = T.ADMMSolver_SPI(den)
    assert torch.equal(spi.get_output(torch.arange(12.).reshape(1, 3, 2, 2)), torch.arange(4.).reshape(1, 1, 2, 2))


def test_factories_and_errors():
    den = T.UNetDenoiser2D(state_dict=weights("he"))
    opt = types.SimpleNamespace(solver="admm", denoiser="unet")
    assert isinstance(T.create_solver_csmri(opt, den), T.ADMMSolver_CSMRI)
    opt.solver = "iadmm"
    assert isinstance(T.create_solver_pr(opt, den), T.IADMMSolver_PR)
    assert isinstance(T.create_solver_ct(opt, den), T.IADMMSolver_CT)
    opt.solver = "admm_spi"
    assert isinstance(T.create_solver_spi(opt, den), T.ADMMSolver_SPI)
    for name, cls in (("hqs", T.HQSSolver_CSMRI), ("pg", T.PGSolver_CSMRI), ("apg", T.APGSolver_CSMRI),
                      ("redadmm", T.REDADMMSolver_CSMRI)):
        opt.solver = name
        assert isinstance(T.create_solver_csmri(opt, den), cls)
    opt.solver = "pg"
    assert isinstance(T.create_solver_ct(opt, den), T.PGSolver_CT)
    opt.solver = "amp"                   # draws random numbers inside the loop: not built -> same error as an unknown name
    with pytest.raises(NotImplementedError):
        T.create_solver_csmri(opt, den)
    opt.denoiser = "dncnn"               # unknown names raise as upstream (tfpnp/pnp/__init__.py:8-13)
    with pytest.raises(NotImplementedError):
        T.create_denoiser(opt, state_dict=weights("he"))
    with pytest.raises(TypeError):
        T.ADMMSolver_CSMRI(torch.nn.Identity())


def test_no_cpu_fallback():
    """CPU tensors must fail loudly, never route through PyTorch."""
    den = T.UNetDenoiser2D(state_dict=weights("he"))
    s = T.ADMMSolver_CSMRI(den)
    st = torch.zeros(1, 3, 32, 32, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        s((st, (torch.zeros(1, 1, 32, 32, 2), torch.zeros(1, 1, 32, 32, dtype=torch.bool))),
          (torch.zeros(1, 2), torch.zeros(1, 2)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        den(torch.zeros(1, 1, 32, 32), torch.zeros(1))


def test_product_never_imports_oracle():
    import glob
    for f in glob.glob(os.path.join(ROOT, "tfpnp_b200", "*.py")):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+.*oracle", src, re.M), f
        assert "oracle" not in src, f


def test_shard_bounds_cover_batch():
    for n in (1, 7, 32, 48, 384):
        for world in (1, 2, 3, 4, 8):
            spans = [T.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    d = {"a": torch.arange(10), "b": [torch.arange(10) * 2, "keep"]}
    s = T.shard_batch(d, 1, 2)
    assert torch.equal(s["a"], torch.arange(5, 10)) and s["b"][1] == "keep"


_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import tfpnp_b200 as T
from oracle import pnp_oracle as O, synth
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
B = 5
g = torch.Generator().manual_seed(0)
out = torch.rand(B, 1, 8, 8, generator=g); gt = torch.rand(B, 1, 8, 8, generator=g)
full = O.psnr(out, gt)                              # what one process would compute
lo, hi = T.shard_bounds(B, rank, 2)
mine = O.psnr(out[lo:hi], gt[lo:hi])                # rank-local metric (oracle stands in for the GPU kernel)
gathered = T.all_gather_psnr(mine, B)
assert gathered.shape == (B, 1), gathered.shape
assert torch.equal(gathered, full), (gathered, full)
imgs = T.all_gather_batch(out[lo:hi], B)               # optional image gather (uneven shards: 3 + 2)
assert torch.equal(imgs, out)
# sharded batches are bit-identical slices
d = synth.spi_batch(4, 16, 1)
sh = T.shard_batch(d, rank, 2)
assert torch.equal(sh["x0"], d["x0"][rank * 2:(rank + 1) * 2])
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_gloo_world2_psnr_gather(tmp_path):
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_pdl_kernels_read_dependent_data_after_the_wait():
    """tools/check_pdl_sass.py: no kernel that executes griddepcontrol.wait loads through a const __restrict__ pointer
    (ld.global.nc, which nvcc may hoist) before it -- the bug csmri_rows_inv had in round 2."""
    import shutil, subprocess, sys
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_pdl_sass.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "kernels execute griddepcontrol.wait" in r.stdout


def test_bench_reference_arm_prints_exactly_one_json_line():
    """bench.py --impl reference (the arm that needs no GPU): stdout is ONE line of JSON with the contract's keys, whatever the
    libraries print while it runs (bench.py sends everything else on fd 1 to stderr)."""
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout[:500]
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_profile_traffic_tag_matches_the_hash_tool():
    """profiles/r02_traffic_*.json carry tools/src_hash.py's hash of the sources they were captured from; bench.py prints
    roofline.traffic only on a match, so a stale capture can never be reported as measured."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from src_hash import src_sha16, CSMRI_ITERATION
    for f in CSMRI_ITERATION:
        assert os.path.exists(os.path.join(ROOT, "tfpnp_b200", "csrc", f)), f
    h = src_sha16(ROOT)
    assert re.fullmatch(r"[0-9a-f]{16}", h) and h == src_sha16(ROOT)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    for prec in ("fp16x3", "fp16"):
        p = os.path.join(ROOT, "profiles", f"r02_traffic_csmri_{prec}.json")
        d = json.load(open(p))
        got = bench.ncu_traffic("csmri", prec)
        assert (got is not None) == (d.get("src_sha16") == h)      # printed only while the tree still hashes to the capture


def test_fft256_half_warp_transform_on_the_host(tmp_path):
    """fft256.cuh (radix-4^2 16-point DFT in registers, lane twiddle, one exchange, second DFT): its __host__ __device__ bodies run
    lane by lane on the CPU against a double-precision DFT, forward and inverse (tests/fft256_host.cu)."""
    import shutil
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("nvcc not available")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "fft256_host")
    r = subprocess.run([nvcc, "-O1", "-Wno-deprecated-gpu-targets", "-I", os.path.join(ROOT, "tfpnp_b200", "csrc"), "-o", exe,
                        os.path.join(ROOT, "tests", "fft256_host.cu")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("max err") == 2


@pytest.mark.parametrize("N,views,scale", [(32, 7, 1.0), (40, 20, 1.0), (64, 24, 1.25), (128, 60, 1.0), (256, 60, 1.0), (48, 181, 0.8)])
def test_ct_ray_walk_clipping_never_drops_a_contributing_step(N, views, scale):
    """ct.cu's projector walks only the driving indices [k_lo, k_hi] where the ray can touch the image.  Restated here in numpy
    fp32 with the kernel's formulas: every step whose interpolation has an in-bounds tap (what the un-clipped, guarded walk
    would have added) lies inside the clipped range -- for the reference's angle table (incl. exactly 0 and ~90 degrees, where
    the slope term vanishes or is ~1e-8), a dense table and tables that are not unit vectors."""
    import numpy as np
    f = np.float32
    ang = np.linspace(0, 179 / 180 * np.pi, views, dtype=np.float32)
    cs = (np.cos(ang.astype(np.float64)).astype(f) * f(scale)).astype(f)
    sn = (np.sin(ang.astype(np.float64)).astype(f) * f(scale)).astype(f)
    D = int(np.ceil(np.sqrt(2) * N))
    c = f((N - 1) * 0.5)
    k = np.arange(N, dtype=np.float32)
    t = (k - c).astype(f)
    for v in range(views):
        co, si = cs[v], sn[v]
        col = abs(si) >= abs(co)
        a, bq = (co, si) if col else (si, co)
        inv_b = f(1.0) / bq
        s = (np.arange(D, dtype=np.float32) - f((D - 1) * 0.5)).astype(f)[:, None]
        r = (((s - t[None, :] * a).astype(f) * inv_b).astype(f) + c).astype(f)        # [D, N]
        i0 = np.floor(r)
        used = (i0 >= -1) & (i0 <= N - 1)                                             # a tap of this step is inside the image
        if a != 0:
            ka = (c + ((s - (f(-1.0) - c) * bq).astype(f) / a).astype(f)).astype(f)
            kb = (c + ((s - (f(N) - c) * bq).astype(f) / a).astype(f)).astype(f)
            lo = np.minimum(np.maximum(np.minimum(ka, kb), f(-2)), f(N + 1))
            hi = np.minimum(np.maximum(np.maximum(ka, kb), f(-2)), f(N + 1))
            k_lo = np.maximum(0, np.floor(lo).astype(np.int64) - 2)
            k_hi = np.minimum(N - 1, np.ceil(hi).astype(np.int64) + 2)
        else:
            k_lo = np.zeros((D, 1), np.int64)
            k_hi = np.full((D, 1), N - 1, np.int64)
        kk = np.arange(N)[None, :]
        inside = (kk >= k_lo) & (kk <= k_hi)
        assert not (used & ~inside).any(), (N, views, scale, v)
        # and the clipping does cut work for oblique views: some steps are skipped somewhere


@pytest.mark.parametrize("N,views", [(32, 9), (64, 24), (128, 60), (256, 60), (256, 180)])
def test_ct_backprojection_windows_cover_every_pixel(N, views):
    """ct.cu's windowed back-projection (unit-norm tables, D = ceil(sqrt(2) N), 16x16 pixel tiles): every pixel's two detector
    bins are inside the detector (0 <= floor(d*) <= D - 2, so the guards can go) and inside the 32-bin window its tile stages
    (window start = clamp(floor(min over the tile's corners of d*) - 1, 0, D - 32)).  numpy fp32 with the kernel's formulas."""
    import numpy as np
    f = np.float32
    ang = np.linspace(0, 179 / 180 * np.pi, views, dtype=np.float32)
    cs = np.cos(ang.astype(np.float64)).astype(f)
    sn = np.sin(ang.astype(np.float64)).astype(f)
    D = int(np.ceil(np.sqrt(2) * N))
    c = f((N - 1) * 0.5)
    half = f((D - 1) * 0.5)
    xs = (np.arange(N, dtype=np.float32) - c).astype(f)
    X, Y = np.meshgrid(xs, xs)                                  # X[i, j] = x of column j, Y[i, j] = y of row i
    for v in range(views):
        co, si = cs[v], sn[v]
        dstar = (((X * co).astype(f) + (Y * si).astype(f)).astype(f) + half).astype(f)
        fl = np.floor(dstar)
        assert fl.min() >= 0 and fl.max() <= D - 2, (N, v, fl.min(), fl.max())
        for ty in range(N // 16):
            for tx in range(N // 16):
                x0, y0 = f(tx * 16) - c, f(ty * 16) - c
                e = [x0 * co + y0 * si, (x0 + f(15)) * co + y0 * si, x0 * co + (y0 + f(15)) * si, (x0 + f(15)) * co + (y0 + f(15)) * si]
                dmin = f(min(e)) + half
                start = min(max(int(np.floor(dmin)) - 1, 0), D - 32)
                o = fl[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16] - start
                assert o.min() >= 0 and o.max() <= 30, (N, v, ty, tx, o.min(), o.max())
