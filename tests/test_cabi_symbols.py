"""CPU: the C-ABI library is built in-tree, loads, and exports every symbol the header declares.
No compute call is made here (there is no GPU in the build container)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from vsum_b200 import _cabi

HEADER = os.path.join(ROOT, "include", "vsum_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(vsum_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(_cabi.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported():
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 14
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/vsum_b200.h but not exported"
    assert sorted(_cabi.EXPORTS) == names


def test_abi_version_and_error_text():
    L = _cabi.load()
    assert L.vsum_abi_version() == 1
    assert L.vsum_knapsack_scratch_words(3, 64) == 3 * 8       # 65 capacities -> class width 256 -> 8 words per shot
    assert L.vsum_knapsack_scratch_words(0, 10) == 0
    import numpy as np
    from vsum_b200.evaluation import _engine
    caps = np.array([0, 1, 254, 255, 256, 1023, 1024, 4095, 4096, 9727, 9728, 18432, 18943, 18944, 28671])
    assert [L.vsum_knapsack_class_width(int(c)) for c in caps] == _engine.knapsack_class_width(caps).tolist()
    assert L.vsum_knapsack_class_width(28672) == -1


def test_no_cpu_fallback_without_cuda():
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from vsum_b200.evaluation import knapSack
    from vsum_b200.model import SimNet
    with pytest.raises(_cabi.VsumError):
        knapSack(7, [2, 2, 1], [4.0, 4.0, 2.0], 3)
    m = SimNet(num_heads=4, d_model=256, num_layers=1).eval()
    with pytest.raises(_cabi.VsumError):
        m(torch.zeros(1, 8, 1024))
