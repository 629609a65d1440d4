"""GPU parity of kernel temporal segmentation (vsum_kts_dp / vsum_kts_gram) against the reference's outputs
(golden fixture) and against the CPU restatement at sizes the reference's Python loops cannot reach."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import kts_ref
from vsum_b200.segmentation import cpd_nonlin, kts_seg, kts_segmentation

pytestmark = pytest.mark.gpu


def same_bits(a, b):
    return np.array_equal(np.asarray(a, np.float64).view(np.int64), np.asarray(b, np.float64).view(np.int64))


def test_golden_from_reference():
    g = np.load(os.path.join(GOLDEN, "kts_golden.npz"))
    for seed, n, dim, ncp, lmin, lmax in g["cases"]:
        K = g[f"K_{seed}"]
        cps, costs = kts_segmentation(K, int(ncp), 1.0, lmin=int(lmin), lmax=int(lmax))
        assert np.array_equal(cps, g[f"cps_{seed}"]), seed
        assert same_bits(costs, g[f"costs_{seed}"]), seed                 # objective values: bit-exact
        cps2, scores = cpd_nonlin(K, int(ncp), lmin=int(lmin), lmax=int(lmax), verbose=False)
        assert np.array_equal(cps2, g[f"cpsfixed_{seed}"]) and same_bits(scores, g[f"scores_{seed}"])
        _, s3 = cpd_nonlin(K, int(ncp), lmin=int(lmin), lmax=int(lmax), backtrack=False)
        assert same_bits(s3, scores)


@pytest.mark.parametrize("n,dim,ncp,kw", [(700, 64, 60, {}), (513, 32, 25, dict(lmin=4, lmax=90)), (1, 8, 0, {}), (2, 8, 1, {}),
                                          (1200, 128, 100, {})])
def test_against_cpu_restatement(n, dim, ncp, kw):
    rng = np.random.default_rng(n)
    x = rng.random((n, dim), dtype=np.float32)
    for c in range(0, n, max(n // 9, 1)):
        x[c:] += rng.random(dim, dtype=np.float32)
    K = np.dot(x, x.T)
    want_cps, want_costs = kts_ref.kts_segmentation(K, ncp, 1.0, **kw)
    cps, costs = kts_segmentation(K, ncp, 1.0, **kw)
    assert np.array_equal(cps, want_cps) and same_bits(costs, want_costs)


def test_assertions_like_the_reference():
    K = np.eye(10, dtype=np.float32)
    with pytest.raises(AssertionError):
        cpd_nonlin(K, 10)                                   # n >= (m + 1) * lmin   (cpd_nonlin.py:46)
    with pytest.raises(AssertionError):
        cpd_nonlin(K, 1, lmin=1, lmax=4)                    # n <= (m + 1) * lmax
    with pytest.raises(AssertionError):
        cpd_nonlin(np.zeros((3, 4), np.float32), 1)


def test_kts_seg_from_features():
    """create_segments.py:24-52: the dot-product kernel on the GPU (fp32 SIMT GEMM) then the same programme; exact against the
    restatement run on the GPU's own K, and the fp32 K itself within 1e-5 of numpy's."""
    from vsum_b200 import _cabi
    rng = np.random.default_rng(7)
    x = rng.random((400, 1024), dtype=np.float32)
    for c in (50, 130, 131, 260, 333):
        x[c:] += 0.3 * rng.random(1024, dtype=np.float32)
    xd = torch.from_numpy(x).cuda()
    K = torch.empty((400, 400), device="cuda")
    zeros = torch.empty(400, device="cuda")
    _cabi.check(_cabi.load().vsum_kts_gram(xd.data_ptr(), 400, 1024, zeros.data_ptr(), K.data_ptr(), torch.cuda.current_stream().cuda_stream), "gram")
    Kh = K.cpu().numpy()
    np.testing.assert_allclose(Kh, np.dot(x, x.T), rtol=1e-5)
    want, _ = kts_ref.kts_segmentation(Kh, 30, 1.0)
    got = kts_seg(x, 30, 1.0)
    assert np.array_equal(got, want) and len(got) > 0
