import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100) device; run with -m gpu on the B200 box")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.dtype.kind in "fbiu" and a.ndim > 0 else a
    return out


def rel_err(a: torch.Tensor, b: torch.Tensor):
    """(relative L2, max-abs / max|ref|) of a against the reference b, in fp64."""
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item(), \
           ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


_WEIGHTS = {}


def weights(init, seed=0):
    from oracle import synth
    key = (init, seed)
    if key not in _WEIGHTS:
        _WEIGHTS[key] = synth.unet_state_dict(seed, init)
    return _WEIGHTS[key]


@pytest.fixture(scope="session", autouse=True)
def _threads():
    torch.set_num_threads(max(1, (os.cpu_count() or 2)))
