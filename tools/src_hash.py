#!/usr/bin/env python
"""16-hex-digit hash of the CUDA sources the library is built from (csrc/*.cu, csrc/*.cuh, include/tfpnp_b200.h).

nvcc's output is not bit-reproducible (two `make -B` runs of the same tree differ), so the ncu captures under profiles/ are tied
to the SOURCE state they were taken from: tools/gpu_profiles.sh records this hash next to them and bench.py prints
`roofline.traffic` only while the running tree still hashes to it."""
import glob, hashlib, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def src_sha16(root=ROOT):
    h = hashlib.sha256()
    files = sorted(glob.glob(os.path.join(root, "tfpnp_b200", "csrc", "*.cu")) + glob.glob(os.path.join(root, "tfpnp_b200", "csrc", "*.cuh")))
    files.append(os.path.join(root, "include", "tfpnp_b200.h"))
    for f in files:
        h.update(os.path.basename(f).encode() + b"\0")
        h.update(open(f, "rb").read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(src_sha16())
