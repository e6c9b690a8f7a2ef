"""pytest configuration: marker registration and shared helpers."""
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, "golden_%s.npz" % name)) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """|a-b|_2 / |b|_2 (0 if both are 0)."""
    a = torch.as_tensor(np.asarray(a)).double().flatten() if not torch.is_tensor(a) else a.detach().double().cpu().flatten()
    b = torch.as_tensor(np.asarray(b)).double().flatten() if not torch.is_tensor(b) else b.detach().double().cpu().flatten()
    den = b.norm().item()
    num = (a - b).norm().item()
    if den == 0:
        return num
    return num / den
