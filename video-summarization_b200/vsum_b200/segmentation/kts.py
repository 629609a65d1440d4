"""`cpd_nonlin`, `kts_segmentation` and `kts_seg` with the reference's names, arguments, assertions and
return values (`cpd_nonlin.py:27-91`, `cpd_auto.py:5-44`, `create_segments.py:24-52`).

The O(n^2) scatter matrix and the O(m n^2) dynamic programme -- pure Python loops in the reference -- run on
the GPU and return the objective row `I[:, n]` plus the back-pointer table; the handful of fp64 operations of
`cpd_auto.py:30-38` (penalties, argmin) and the back-tracking run here on those tables.  Because the table
rows for k change points do not depend on the total number requested, the reference's second
`cpd_nonlin(K, m_best)` pass is a back-track over the tables of the first one."""
from __future__ import annotations

import numpy as np
import torch

from .. import _cabi


def _device():
    if not torch.cuda.is_available():
        raise _cabi.VsumError("vsum_b200.segmentation needs a CUDA device: the B200 kernels have no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _tables(K, m: int, lmin: int, lmax: int):
    """-> (scores fp64[m+1] = I[:, n], p int32[m+1, n+1]) as numpy arrays."""
    dev = _device()
    Kd = K if isinstance(K, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(K), dtype=np.float32))
    Kd = Kd.to(dev, dtype=torch.float32).contiguous()
    n = Kd.shape[0]
    L = _cabi.load()
    with torch.cuda.device(dev):
        need = L.vsum_kts_workspace_bytes(n, m)
        ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
        wp = (ws.data_ptr() + 1023) // 1024 * 1024
        scores = torch.empty(m + 1, dtype=torch.float64, device=dev)
        prev = torch.empty((m + 1, n + 1), dtype=torch.int32, device=dev)
        _cabi.check(L.vsum_kts_dp(Kd.data_ptr(), n, m, int(lmin), int(min(lmax, 2 ** 31 - 1)), wp, ws.numel() - (wp - ws.data_ptr()),
                                  scores.data_ptr(), prev.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "vsum_kts_dp")
        return scores.cpu().numpy(), prev.cpu().numpy()


def _checks(K, ncp, lmin, lmax):
    m = int(ncp)
    n, n1 = K.shape
    assert n == n1, "Kernel matrix awaited."
    assert n >= (m + 1) * lmin
    assert n <= (m + 1) * lmax
    assert lmax >= lmin >= 1
    return m, n


def _backtrack(prev, m, n):
    cps = np.zeros(m, dtype=int)
    cur = n
    for k in range(m, 0, -1):
        cps[k - 1] = prev[k, cur]
        cur = cps[k - 1]
    return cps


def cpd_nonlin(K, ncp, lmin=1, lmax=100000, backtrack=True, verbose=True, out_scatters=None):
    """Change-point detection for a fixed number of change points -> (cps, scores).  `out_scatters` is not
    supported (the scatter matrix stays on the device)."""
    if out_scatters is not None:
        raise NotImplementedError("out_scatters: the scatter matrix is not copied back from the device")
    m, n = _checks(K, ncp, lmin, lmax)
    scores, prev = _tables(K, m, lmin, lmax)
    cps = _backtrack(prev, m, n) if backtrack else np.zeros(m, dtype=int)
    scores = scores.copy()
    scores[scores > 1e99] = np.inf
    return cps, scores


def kts_segmentation(K, ncp, vmax, desc_rate=1, **kwargs):
    """Automatic selection of the number of change points -> (cps, costs) (cpd_auto.py:5-44)."""
    lmin, lmax = kwargs.get("lmin", 1), kwargs.get("lmax", 100000)
    m, n = _checks(K, ncp, lmin, lmax)
    scores, prev = _tables(K, m, lmin, lmax)
    scores = scores.copy()
    scores[scores > 1e99] = np.inf
    N = n
    N2 = N * desc_rate
    penalties = np.zeros(m + 1)
    ncp_r = np.arange(1, m + 1)
    penalties[1:] = (vmax * ncp_r / (2.0 * N2)) * (np.log(float(N2) / ncp_r) + 1)
    costs = scores / float(N) + penalties
    m_best = int(np.argmin(costs))
    _checks(K, m_best, lmin, lmax)                      # the reference's second cpd_nonlin call asserts again
    return _backtrack(prev, m_best, n), costs


def kts_seg(features, num_seg: int, v_max: float, kernel: str = "dot"):
    """create_segments.py:24-52: change points of a video from its frame features [n, dim] (dot-product kernel
    computed on the GPU in fp32)."""
    if kernel != "dot":
        raise NotImplementedError
    dev = _device()
    x = features if isinstance(features, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(features), dtype=np.float32))
    x = x.to(dev, dtype=torch.float32).contiguous()
    n, dim = x.shape
    with torch.cuda.device(dev):
        K = torch.empty((n, n), dtype=torch.float32, device=dev)
        zeros = torch.empty(n, dtype=torch.float32, device=dev)
        _cabi.check(_cabi.load().vsum_kts_gram(x.data_ptr(), n, dim, zeros.data_ptr(), K.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "vsum_kts_gram")
    segments, _ = kts_segmentation(K, num_seg, v_max)
    return segments
