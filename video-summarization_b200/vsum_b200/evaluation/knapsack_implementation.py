"""`knapSack` with the reference signature (`src/evaluation/knapsack_implementation.py:1-30`),
solved by the shared-memory DP kernel behind `vsum_knapsack` (one video = one CTA)."""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi
from . import _engine


def knapSack(W, wt, val, n):
    """Capacity W (int), weights `wt`, values `val`, item count n -> ascending list of chosen
    item indices.  Values are taken as Python floats (fp64); ties go to the lower index."""
    n, W = int(n), int(W)
    if n <= 0 or W < 0:
        return []
    dev = _engine._device()
    L = _cabi.load()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        d_val = torch.tensor([float(v) for v in list(val)[:n]], dtype=torch.float64, device=dev)
        d_wt = torch.tensor([int(w) for w in list(wt)[:n]], dtype=torch.int32, device=dev)
        d_cu = torch.tensor([0, n], dtype=torch.int32, device=dev)
        d_cap = torch.tensor([W], dtype=torch.int32, device=dev)
        words = int(L.vsum_knapsack_scratch_words(n, W))
        if words < 0:
            raise _cabi.VsumError(f"knapSack: capacity {W} exceeds the largest sm_100a kernel class (28671)")
        d_off = torch.tensor([0, words], dtype=torch.int64, device=dev)
        bits = torch.empty(max(words, 1), dtype=torch.int32, device=dev)
        sel = torch.empty(n, dtype=torch.uint8, device=dev)
        _cabi.check(L.vsum_knapsack(d_val.data_ptr(), d_wt.data_ptr(), d_cu.data_ptr(), d_cap.data_ptr(),
                                    d_off.data_ptr(), None, 1, W, bits.data_ptr(), sel.data_ptr(), stream),
                    "vsum_knapsack")
        return [int(i) for i in np.nonzero(sel.cpu().numpy())[0]]
