#!/usr/bin/env python
"""Headline benchmark: PnP-ADMM inner-iterations per second (image-iterations/s).

Workload (BASELINE.json configs[1]): CS-MRI ADMM, env_batch=48, 128x128, action_pack=5 x
max_episode_step=6 = 30 inner iterations per solver call, UNet denoiser, synthetic k-space
batches (seed 1234), seeded default-init UNet(2,1) weights.
One "step" = one `solver(inputs, parameters)` call = 48 images x 30 iterations per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision fp16|fp16x3] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); the batch dimension is sharded with no
data-path collective (weak scaling: 48 images per GPU) and one NCCL all-gather of the PSNR vector
per step.  `--impl reference` times the reference algorithm on the host CPU (the oracle port of the
reference's PyTorch path; the reference itself is 100 % Python and cannot travel to the GPU box).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

B_PER_GPU, N_PIX, ITERS = 48, 128, 30
GFLOP_PER_IMAGE_ITER = 9.6836          # UNet(2,1) at 128x128, 2*MAC
UPDATE_BYTES_PER_PX = 37 + 4           # CS-MRI fused update: SURVEY 8d figure + the d = Re(z-u) write
METRIC = "PnP-ADMM inner-iters/sec (env_batch x H x W images/sec)"
UNIT = "image-iters/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    fb = dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            hbm = d.get("hbm_gbs")
            burst = d.get("bf16_tflops")
            sust = d.get("bf16_tflops_sustained", burst)
            if hbm and (sust or burst):
                return dict(hbm=float(hbm), tf_burst=float(burst or sust), tf_sust=float(sust or burst), src="measured")
        except Exception:
            pass
    return fb


def ncu_traffic():
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of one inner iteration's kernels from the committed
    `ncu --set full` capture of this same command (profiles/r01_traffic.json, written by tools/ncu_traffic.py)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_inputs(T, dev, B, n, iters, seed):
    """Synthetic CS-MRI batch of the BASELINE shape, built with the PRODUCT's own operators on the GPU:
    gt ~ U[0,1), radial masks cycling over ~50/25/12.5 % sampling, y0 = fft2(gt) + N(0,(15/255)^2) on the mask,
    x0 = ifft2(y0), state = ADMMSolver.reset, sigma_d ~ U[0,70/255], mu ~ U[0,1].  Returned on the HOST (pinned)."""
    g = torch.Generator(dev).manual_seed(seed)
    gt = torch.rand(B, 1, n, n, device=dev, generator=g)
    opts = [max(2, n // 3), max(2, n // 6), max(2, n // 12)]
    masks = [T.radial_mask(n, L, device=dev) for L in opts]
    mask = torch.stack([masks[b % 3] for b in range(B)])[:, None]
    m = T.csmri_measure(gt, mask, sigma_n=15 / 255, generator=g)
    x0 = m["x0"]
    state = torch.cat((x0, x0.clone(), torch.zeros_like(x0)), dim=1)          # ADMMSolver.reset (base.py:95-99)
    sigma_d = torch.rand(B, iters, device=dev, generator=g) * (70 / 255)
    mu = torch.rand(B, iters, device=dev, generator=g)
    d = dict(state=state, y0=m["y0"], mask=m["mask"], sigma_d=sigma_d, mu=mu, gt=gt)
    return {k: v.cpu().contiguous() for k, v in d.items()}


def cpu_reference_rate(sd, d, B, iters, repeats=1):
    """The reference algorithm (oracle port of tasks/csmri/solver.py:29-57 + UNetDenoiser2D) on the host."""
    from oracle import pnp_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    sl = slice(0, B)
    args = (d["state"][sl], d["y0"][sl], d["mask"][sl], d["sigma_d"][sl, :iters], d["mu"][sl, :iters])
    best = float("inf")
    with torch.no_grad():
        for _ in range(repeats):
            t0 = time.perf_counter()
            O.admm_csmri(sd, *args)
            best = min(best, time.perf_counter() - t0)
    return B * iters / best, best


def run_reference(args):
    """`--impl reference`: rank 0 alone times the CPU path; other ranks exit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import synth
    sd = synth.unet_state_dict(0, "default")
    d = synth.csmri_batch(B_PER_GPU, N_PIX, ITERS)
    sample_it = 1                                     # one step = 48 images x 1 iteration (bounded sample)
    for _ in range(args.warmup):
        cpu_reference_rate(sd, d, 8, 1)
    times = []
    for _ in range(args.steps):
        _, t = cpu_reference_rate(sd, d, B_PER_GPU, sample_it)
        times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = B_PER_GPU * sample_it / (ms / 1e3)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "csmri ADMM, env_batch=48, 128x128, UNet denoiser (BASELINE configs[1])",
                   "sample": f"B=48 x {sample_it} inner iteration per step (loop is linear in B*iters)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle port of the reference PyTorch path, B=48, 128x128, {sample_it} iter/step, "
                                   f"{cores} threads of {os.cpu_count()} host CPUs"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fp16", choices=["fp16", "fp16x3", "fp32_simt"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch.distributed as dist
    import tfpnp_b200 as T       # the product arm imports nothing from oracle/ (only the cpu_baseline / parity leg below does)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    # ---- inputs: weak scaling, 48 images per GPU, different seed per rank --------------------
    sd = T.random_unet_state_dict(0)
    d = synth_inputs(T, dev, B_PER_GPU, N_PIX, ITERS, seed=1234 + rank)
    B_total = B_PER_GPU * world
    solver = T.ADMMSolver_CSMRI(T.UNetDenoiser2D(state_dict=sd, precision=args.precision))
    host = {k: d[k].contiguous().pin_memory() for k in ("state", "y0", "mask", "sigma_d", "mu", "gt")}
    res = {k: host[k].to(dev) for k in host}                    # resident copies for `value`
    out_host = torch.empty_like(host["state"]).pin_memory()
    psnr_host = torch.empty(B_total, 1).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_resident():
        with torch.no_grad():
            out = solver((res["state"], (res["y0"], res["mask"])), (res["sigma_d"], res["mu"]))
            p = T.all_gather_psnr(T.torch_psnr(solver.get_output(out), res["gt"]), B_total)
        return out, p

    def step_e2e():
        with torch.no_grad():
            g = {k: host[k].to(dev, non_blocking=True) for k in ("state", "y0", "mask", "sigma_d", "mu", "gt")}
            out = solver((g["state"], (g["y0"], g["mask"])), (g["sigma_d"], g["mu"]))
            p = T.all_gather_psnr(T.torch_psnr(solver.get_output(out), g["gt"]), B_total)
            out_host.copy_(out, non_blocking=True)
            psnr_host.copy_(p, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        evs = []
        barrier()
        for _ in range(steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            evs.append((a, b))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in evs) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    for _ in range(args.warmup):
        step_resident()
    step_e2e()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res = timed(step_resident, args.steps)
    launches = solver.last_launch_count + 1       # + the PSNR kernel
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernel-class breakdown (eager replay with events inside the library) ----------------
    import ctypes as C
    from tfpnp_b200 import _lib
    h = next(iter(solver._solvers.values()))
    _lib.lib().tfpnp_solver_set_profiling(h, 1)
    step_resident(); step_resident()
    den_ms, upd_ms = C.c_float(), C.c_float()
    _lib.check(_lib.lib().tfpnp_solver_get_profile(h, C.byref(den_ms), C.byref(upd_ms)), "get_profile")
    _lib.lib().tfpnp_solver_set_profiling(h, 0)
    pk = peaks()
    den_tflops = B_PER_GPU * ITERS * GFLOP_PER_IMAGE_ITER / den_ms.value           # GFLOP/ms = TFLOP/s
    upd_gbs = B_PER_GPU * N_PIX * N_PIX * UPDATE_BYTES_PER_PX * ITERS / upd_ms.value / 1e6

    value = B_total * ITERS / (ms_res / 1e3)
    e2e = B_total * ITERS / (ms_e2e / 1e3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("state", "y0", "mask", "sigma_d", "mu", "gt"))
    d2h = out_host.numel() * 4 + B_total * 4

    tr = ncu_traffic()
    den_traffic = tr["denoiser_dram_bytes_per_iter"] * ITERS if tr else None       # per step, like `achieved`
    upd_traffic = tr["update_dram_bytes_per_iter"] * ITERS if tr else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16" if args.precision != "fp32_simt" else "f32", "data": "synthetic",
        "config": {"workload": "csmri ADMM, env_batch=48/GPU, 128x128, action_pack=5 x max_episode_step=6 "
                               "(30 inner iters per call), UNet denoiser (BASELINE configs[1])",
                   "precision": args.precision, "global_batch": B_total, "iters_per_step": ITERS,
                   "l2": "256 MiB flush write between timed steps", "weights": "seeded default-init UNet(2,1)",
                   "inputs": "synthesised on the GPU by tfpnp_b200.csmri_measure (radial masks, sigma_n = 15/255)"},
        "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches * args.steps),
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv3x3_tc (tcgen05 implicit-GEMM, denoiser segment)",
                     "achieved": den_tflops, "peak": pk["tf_sust"], "unit": "TFLOP/s",
                     "frac": den_tflops / pk["tf_sust"], "traffic": den_traffic,
                     "traffic_note": "DRAM bytes per step (30 denoiser calls) from profiles/r01_traffic.json (ncu --set full)",
                     "peak_source": pk["src"] + " bf16 sustained",
                     "denoiser_ms_per_step": den_ms.value, "update_ms_per_step": upd_ms.value},
        "roofline_update": {"bound": "hbm", "kernel": "csmri rows_fwd+cols+rows_inv", "achieved": upd_gbs,
                            "peak": pk["hbm"], "unit": "GB/s", "frac": upd_gbs / pk["hbm"], "traffic": upd_traffic,
                            "algorithmic_bytes_per_step": B_PER_GPU * N_PIX * N_PIX * UPDATE_BYTES_PER_PX * ITERS},
    }

    if rank == 0 and not args.no_cpu_baseline:
        # bounded CPU sample of the same workload on this box's host cores (rank 0, N=1 semantics)
        sample_B, sample_it = 48, 2
        rate, secs = cpu_reference_rate(sd, d, sample_B, sample_it)
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"oracle port of the reference PyTorch path: B={sample_B}, 128x128, "
                                          f"{sample_it} iters = {secs:.1f} s on {os.cpu_count()} host CPUs"}
        # parity spot check of the timed configuration against the oracle (2 images, all 30 iterations)
        from oracle import pnp_oracle as O
        with torch.no_grad():
            ref = O.admm_csmri(sd, d["state"][:2], d["y0"][:2], d["mask"][:2], d["sigma_d"][:2], d["mu"][:2])
            out = solver((res["state"][:2], (res["y0"][:2], res["mask"][:2])), (res["sigma_d"][:2], res["mu"][:2]))
        err = ((out.cpu() - ref).abs().max() / ref.abs().max()).item()
        line["parity"] = {"rel_max_err_vs_oracle": err, "images": 2, "iters": ITERS, "tolerance": 1e-4}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
