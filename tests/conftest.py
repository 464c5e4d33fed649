import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """Tests marked gpu are skipped (not failed) when collected on a machine without CUDA and
    without an explicit -m selection; the driver selects them with -m gpu on the GPU box."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
            return {k: z[k] for k in z.files}
    return load


def rel_err(a, b, floor=None):
    """max |a-b| / max(|b|, floor) per element (NaNs must coincide)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert (nan_a == nan_b).all(), "NaN pattern differs"
    ok = ~nan_a
    if not ok.any():
        return 0.0
    scale = np.abs(b[ok])
    if floor is not None:
        scale = np.maximum(scale, floor)
    scale = np.maximum(scale, np.finfo(np.float64).tiny)
    return float(np.max(np.abs(a[ok] - b[ok]) / scale))
