import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (sm_100a) device; run on the B200 box with -m gpu")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: z[k] for k in z.files}


def reverse_noises(g):
    """the Gaussian draws the reference's predict_tsc took for tests/golden/diffusion_reverse_*.npz: torch.manual_seed(noise_seed),
    then one randn_like(audio) of shape (1, Lp) per step n = N-1 .. 1 (inference_diffuse.py:259)."""
    L = g["noisy"].shape[-1]
    Lp = -(-L // 100) * 100
    gen = torch.Generator().manual_seed(int(g["noise_seed"]))
    n_steps = len(g["c1"])
    return {n: torch.randn(1, Lp, generator=gen) for n in range(n_steps - 1, 0, -1)}


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session", autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def rel_max(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a - b| / max |b|"""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))
