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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name, device="cpu"):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a).to(device) if a.ndim > 0 else a.item()
    return out


def assert_close_rel(y, ref, rel=1e-5, what=""):
    """The parity metric fixed in SURVEY §7: max|y - ref| <= rel * max|ref| per output tensor
    (element-wise relative error is meaningless where the reference is ~0, e.g. Q11)."""
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    assert y.shape == ref.shape, f"{what}: shape {tuple(y.shape)} vs {tuple(ref.shape)}"
    scale = ref.abs().max().item() if ref.numel() else 0.0
    err = (y - ref).abs().max().item() if ref.numel() else 0.0
    assert err <= rel * max(scale, 1e-30), f"{what}: max abs err {err:.3e} > {rel:g} * {scale:.3e}"


@pytest.fixture
def golden():
    return load_golden
