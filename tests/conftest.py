"""pytest configuration: markers, import paths and shared fixtures."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "vqa-project_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.fixture(scope="session", params=["tiny", "small"])
def golden(request):
    return request.param, load_golden(request.param)


def golden_params(g, dtype=torch.float32):
    return {k[len("param."):]: torch.from_numpy(v).to(dtype) for k, v in g.items() if k.startswith("param.")}


def rel_err(a, b):
    """max-norm relative error ||a-b||_inf / ||b||_inf (the tolerance BASELINE.json's north_star states)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    den = b.abs().max().item()
    return (a - b).abs().max().item() / (den if den > 0 else 1.0)
