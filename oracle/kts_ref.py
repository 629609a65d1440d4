"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of kernel temporal segmentation as the reference runs it
(`src/data/preprocess/segmentations/kts/cpd_nonlin.py:5-91`, `cpd_auto.py:5-44`): the same arithmetic in
the same order (float32 cumulative sums of K, fp64 scatters, fp64 dynamic programme with the strict `<`
first-minimum rule), with the two innermost Python loops vectorised.  Pinned against the imported
reference in tests/test_oracle_vs_reference.py and through tests/golden/kts_golden.npz.
"""
import numpy as np


def calc_scatters(K):
    """cpd_nonlin.py:5-24.  K1 is fp64 (a Python list starting with int 0), the double cumsum of a float32 K
    stays float32 before it lands in the fp64 K2."""
    n = K.shape[0]
    K1 = np.cumsum([0] + list(np.diag(K)))
    K2 = np.zeros((n + 1, n + 1))
    K2[1:, 1:] = np.cumsum(np.cumsum(K, 0), 1)
    scatters = np.zeros((n, n))
    d2 = np.diag(K2)
    for i in range(n):
        j = np.arange(i, n)
        scatters[i, i:] = K1[j + 1] - K1[i] - (d2[j + 1] + K2[i, i] - K2[j + 1, i] - K2[i, j + 1]) / (j - i + 1)
    return scatters


def cpd_nonlin(K, ncp, lmin=1, lmax=100000, backtrack=True):
    """cpd_nonlin.py:27-91 -> (cps, scores)."""
    m = int(ncp)
    n, n1 = K.shape
    assert n == n1 and n >= (m + 1) * lmin and n <= (m + 1) * lmax and lmax >= lmin >= 1
    J = calc_scatters(K)
    I = 1e101 * np.ones((m + 1, n + 1))
    I[0, lmin:lmax] = J[0, lmin - 1:lmax - 1]
    p = np.zeros((m + 1, n + 1), dtype=int)
    for k in range(1, m + 1):
        for l in range((k + 1) * lmin, n + 1):
            t0, t1 = max(k * lmin, l - lmax), l - lmin          # inclusive candidate range (line 74)
            best, arg = 1e100, 0
            if t1 >= t0:
                c = I[k - 1, t0:t1 + 1] + J[t0:t1 + 1, l - 1]
                a = int(np.argmin(c))                              # first minimum == strict `<` scan (line 76)
                if c[a] < 1e100:
                    best, arg = c[a], t0 + a
            I[k, l], p[k, l] = best, arg
    cps = np.zeros(m, dtype=int)
    if backtrack:
        cur = n
        for k in range(m, 0, -1):
            cps[k - 1] = p[k, cur]
            cur = cps[k - 1]
    scores = I[:, n].copy()
    scores[scores > 1e99] = np.inf
    return cps, scores


def kts_segmentation(K, ncp, vmax, desc_rate=1, **kw):
    """cpd_auto.py:5-44 -> (cps, costs)."""
    m = ncp
    _, scores = cpd_nonlin(K, m, backtrack=False, **kw)
    N = K.shape[0]
    N2 = N * desc_rate
    penalties = np.zeros(m + 1)
    ncp_r = np.arange(1, m + 1)
    penalties[1:] = (vmax * ncp_r / (2.0 * N2)) * (np.log(float(N2) / ncp_r) + 1)
    costs = scores / float(N) + penalties
    m_best = np.argmin(costs)
    cps, _ = cpd_nonlin(K, m_best, **kw)
    return cps, costs
