import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "video-summarization_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference/src"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def eval_golden():
    return np.load(os.path.join(GOLDEN, "eval_golden.npz"))


@pytest.fixture(scope="session")
def scorer_golden():
    return np.load(os.path.join(GOLDEN, "scorer_golden.npz"))


@pytest.fixture(scope="session")
def scorer_long_golden():
    """Reference logits at N = 4096 and 8192 (tests/golden/make_golden.py::scorer_long_goldens)."""
    return np.load(os.path.join(GOLDEN, "scorer_long_golden.npz"))


@pytest.fixture(scope="session")
def seeded_model_kwargs():
    return dict(num_heads=4, d_model=256, num_layers=4, sparsity=0., use_cls=False, dropout=0.3,
                num_classes=1, use_pos=True)


def bits_equal(a, b) -> bool:
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()
