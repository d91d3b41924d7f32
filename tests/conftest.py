"""pytest configuration: the `gpu` marker (tests that need a B200) and shared helpers."""
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def rel_l2(a, b):
    """||a-b|| / ||b|| over all elements (complex ok); 0/0 -> 0."""
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    if a.is_complex() or b.is_complex():
        a, b = torch.view_as_real(a.to(torch.complex128)), torch.view_as_real(b.to(torch.complex128))
    a, b = a.double(), b.double()
    den = b.norm().item()
    num = (a - b).norm().item()
    return num / den if den > 0 else num


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def to_torch(d, dtype=None):
    out = {}
    for k, v in d.items():
        t = torch.from_numpy(np.asarray(v))
        if dtype is not None and t.is_floating_point():
            t = t.to(dtype)
        elif dtype == torch.float64 and t.is_complex():
            t = t.to(torch.complex128)
        out[k] = t
    return out
