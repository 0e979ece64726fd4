#!/usr/bin/env python
"""16-hex-digit hash of the CUDA sources behind the kernels of one captured inner iteration.

nvcc's output is not bit-reproducible (two `make -B` runs of the same tree differ), so the ncu captures under profiles/ are tied
to the SOURCE state they were taken from: tools/gpu_profiles.sh records this hash next to them and bench.py prints
`roofline.traffic` only while the running tree still hashes to it.  The hash covers the files the captured kernels (the CS-MRI
inner iteration: denoiser layers + the three update kernels) and their launch plans are compiled from -- not the other tasks'
kernels or the reverse-mode code, which do not run in that capture."""
import hashlib, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSMRI_ITERATION = ("unet_tc.cu", "conv_x3.cuh", "conv_ws.cuh", "sm100.cuh", "grad_elem.cuh", "csmri.cu", "fft.cuh", "solver.cu")


def src_sha16(root=ROOT, files=CSMRI_ITERATION):
    h = hashlib.sha256()
    for f in sorted(files):
        h.update(f.encode() + b"\0")
        h.update(open(os.path.join(root, "tfpnp_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(src_sha16())
