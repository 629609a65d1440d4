"""GPU parity (bit-exact): shot pooling, knapsack, summary mask and F-score kernels, called
through the C ABI, against the C oracle and the golden fixtures of the reference."""
import numpy as np
import pytest
import torch

from conftest import bits_equal
from oracle import c_oracle
from vsum_b200.evaluation import _engine, eval_fscores, eval_metrics, evaluate_summary, generate_summary, knapSack
from vsum_b200.synthetic import make_scores, make_video

pytestmark = pytest.mark.gpu


def run_batch(videos, scores, method="avg"):
    hb = _engine.HostEvalBatch.build([v.change_points for v in videos], [np.array(v.n_frames) for v in videos],
                                     [v.picks for v in videos], [v.user_summary for v in videos])
    db = _engine.DeviceEvalBatch(hb)
    cu = torch.from_numpy(_engine._cu([len(s) for s in scores]).astype(np.int32)).to(db.device)
    out = _engine.summarize(db, torch.from_numpy(np.concatenate(scores)).to(db.device), cu, method, want_per_user=True)
    torch.cuda.synchronize()
    return hb, {k: (v.cpu().numpy() if v is not None else None) for k, v in out.items()}


def check_against_oracle(videos, scores, method):
    hb, out = run_batch(videos, scores, method)
    for i, (v, sc) in enumerate(zip(videos, scores)):
        want = c_oracle.video(sc, v.picks, v.n_frames, v.change_points, v.user_summary, method)
        s0, s1 = hb.cu_shots[i], hb.cu_shots[i + 1]
        assert bits_equal(out["val"][s0:s1], want["val"]), f"shot means differ, video {i} (N={v.n_steps})"
        assert bits_equal(out["wt"][s0:s1], want["wt"])
        assert int(out["cap"][i]) == want["cap"]
        assert bits_equal(out["selected"][s0:s1], want["selected"]), f"knapsack selection differs, video {i} (N={v.n_steps})"
        assert bits_equal(out["summary"][hb.sum_offsets[i]:hb.sum_offsets[i + 1]], want["summary"])
        assert bits_equal(np.float64(out["f"][i]), np.float64(want["f"])), f"F differs, video {i}"


def test_golden_videos_bit_exact(eval_golden):
    cases = [tuple(int(x) for x in r) for r in eval_golden["cases"]]
    videos = [make_video(vid, n, n_users=u, with_features=False) for vid, n, u in cases]
    scores = [make_scores(vid, n) for vid, n, _ in cases]
    for k, method in enumerate(("avg", "max")):
        hb, out = run_batch(videos, scores, method)
        for i, (vid, n, u) in enumerate(cases):
            summ = out["summary"][hb.sum_offsets[i]:hb.sum_offsets[i + 1]]
            want = np.unpackbits(eval_golden[f"summary_{vid}"])[:len(summ)].astype(np.int8)
            assert bits_equal(summ, want), (vid, n)
            assert bits_equal(out["val"][hb.cu_shots[i]:hb.cu_shots[i + 1]], eval_golden[f"means_{vid}"]), (vid, n)
            assert bits_equal(np.float64(out["f"][i]), np.float64(eval_golden[f"f_{vid}"][k])), (vid, n, method)


@pytest.mark.parametrize("method", ["avg", "max"])
def test_random_batch_against_oracle(method):
    rng = np.random.default_rng(21)
    ns = [int(x) for x in np.exp(rng.uniform(np.log(1), np.log(3000), 160))]
    videos = [make_video(1000 + i, n, n_users=int(rng.integers(1, 21)), with_features=False) for i, n in enumerate(ns)]
    check_against_oracle(videos, [make_scores(1000 + i, n) for i, n in enumerate(ns)], method)


def test_full_size_videos_against_oracle():
    # BASELINE config 5 upper end: N = 8192 (S ~ 819 shots, capacity ~ 18.4 k frames)
    ns = [8192, 8192, 6000, 4096, 128]
    videos = [make_video(2000 + i, n, n_users=20, with_features=False) for i, n in enumerate(ns)]
    check_against_oracle(videos, [make_scores(2000 + i, n) for i, n in enumerate(ns)], "avg")


def test_ties_and_degenerate_scores():
    # constant scores make every shot's value equal: exercises the tie rule (line 26) at scale
    videos = [make_video(3000 + i, n, n_users=5, with_features=False) for i, n in enumerate([300, 1000, 2500])]
    for const in (0.5, 1.0, 0.0):
        check_against_oracle(videos, [np.full(v.n_steps, const, np.float32) for v in videos], "avg")
    # two-level scores: many exact ties between shots of different lengths
    rng = np.random.default_rng(1)
    scores = [rng.choice(np.array([0.25, 0.75], np.float32), v.n_steps) for v in videos]
    check_against_oracle(videos, scores, "max")


def test_nothing_fits_gives_nan():
    cps = np.array([[0, 99]], dtype=np.int32)
    us = np.ones((2, 100), dtype=np.float32)
    hb = _engine.HostEvalBatch.build([cps], [np.array(100)], [np.arange(0, 100, 15, dtype=np.int32)], [us])
    db = _engine.DeviceEvalBatch(hb)
    out = _engine.summarize(db, torch.full((7,), 0.5, device=db.device), torch.tensor([0, 7], dtype=torch.int32, device=db.device))
    assert out["summary"].sum().item() == 0 and np.isnan(out["f"].cpu().numpy()[0])


def test_reference_call_surface(eval_golden):
    assert knapSack(7, [2, 2, 1, 1, 1, 2], [4, 4, 2, 2, 2, 4], 6) == [0, 1, 2, 3, 4]   # knapsack_implementation.py:35-41
    assert knapSack(2, [2, 1, 1], [2, 1, 1], 3) == [0] and knapSack(2, [1, 1, 2], [1, 1, 2], 3) == [0, 1]
    assert knapSack(0, [1], [1.0], 1) == []
    rng = np.random.default_rng(2)
    for _ in range(20):
        n = int(rng.integers(1, 40))
        wt, val, W = rng.integers(1, 50, n), rng.choice(rng.random(6), n), int(rng.integers(0, 200))
        assert knapSack(W, wt, [float(x) for x in val], n) == c_oracle.knapsack(W, wt, val)
    v = make_video(6, 300, n_users=20, with_features=False)
    sc = make_scores(6, 300)
    summ = generate_summary([v.change_points], [sc], [np.array(v.n_frames)], [v.picks])
    assert len(summ) == 1 and summ[0].dtype == np.int8
    assert bits_equal(summ[0], np.unpackbits(eval_golden["summary_6"])[:len(summ[0])].astype(np.int8))
    for k, method in enumerate(("avg", "max")):
        assert bits_equal(np.float64(evaluate_summary(summ[0], v.user_summary, method)), np.float64(eval_golden["f_6"][k]))
    # summary shorter / longer than the user matrix (evaluation_metrics.py:12-15)
    short, us = np.array([1, 0, 1], np.int8), np.array([[1, 1, 0, 1, 0]], np.float32)
    assert bits_equal(np.float64(evaluate_summary(short, us, "avg")), np.float64(c_oracle.fscore(short, us, "avg")[0]))
    long_ = np.ones(9, np.int8)
    assert bits_equal(np.float64(evaluate_summary(long_, us, "max")), np.float64(c_oracle.fscore(long_, us, "max")[0]))


def test_fscore_uint8_user_summaries_match_float32():
    """The packed dataset's uint8 user summaries go through their own overlap kernel: same counts, same fp64 F as the
    float32 rows, for every alignment of the rows and of the mask, ragged lengths and values other than 0 / 1."""
    rng = np.random.default_rng(77)
    masks, users_f32 = [], []
    for k in range(40):
        cols = int(rng.integers(1, 700)) if k % 4 else int(rng.integers(3000, 9000))
        slen = max(1, cols + int(rng.integers(-40, 40))) if k % 3 else cols
        U = int(rng.integers(1, 6))
        us = (rng.random((U, cols)) < 0.3).astype(np.float32)
        if k % 5 == 0:
            us[rng.integers(0, U), rng.integers(0, cols)] = float(rng.integers(2, 256))     # general (non 0/1) path
        us[:, 0] = 1.0                                     # no 0/0: NaN payloads are not part of the contract
        m = (rng.random(slen) < 0.2).astype(np.int8)
        m[0] = 1
        masks.append(m)
        users_f32.append(us)
    for method in ("avg", "max"):
        f32 = _engine.fscore_of_masks(masks, users_f32, method)
        u8 = _engine.fscore_of_masks(masks, [u.astype(np.uint8) for u in users_f32], method)
        assert bits_equal(f32, u8)
        for m, us, f in zip(masks[:12], users_f32[:12], u8[:12]):
            assert bits_equal(np.float64(f), np.float64(c_oracle.fscore(m, us, method)[0]))
    # whole batched path: uint8 rows in the evaluation batch
    videos = [make_video(v, n, n_users=7, with_features=False) for v, n in ((3, 150), (4, 333), (5, 64), (9, 901))]
    scores = [make_scores(v, n) for v, n in ((3, 150), (4, 333), (5, 64), (9, 901))]
    dev_scores = torch.from_numpy(np.concatenate(scores)).cuda()
    cu = torch.from_numpy(_engine._cu([len(s) for s in scores]).astype(np.int32)).cuda()
    out = []
    for cast in (np.float32, np.uint8):
        hb = _engine.HostEvalBatch.build([v.change_points for v in videos], [v.n_frames for v in videos],
                                         [v.picks for v in videos], [v.user_summary.astype(cast) for v in videos])
        assert hb.user_summary.dtype == cast
        out.append(_engine.summarize(_engine.DeviceEvalBatch(hb), dev_scores, cu, "avg")["f"].cpu().numpy())
    assert bits_equal(out[0], out[1])


def test_eval_metrics_golden():
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "eval_metrics_golden.npz"))
    data, users = {}, {}
    for vid, n in zip(g["ids"], g["ns"]):
        v = make_video(int(vid), int(n), n_users=20, with_features=False, with_user_scores=True)
        data[v.name], users[v.name] = make_scores(int(vid), int(n)), v.as_user()
    f, tau, rho = eval_metrics(data, users)
    assert bits_equal(np.float64(f), np.float64(g["result"][0]))          # F: bit-exact
    assert bits_equal(np.float64(tau), np.float64(g["result"][1]))        # Kendall tau on the GPU: bit-exact with the reference
    assert abs(rho - g["result"][2]) < 1e-12                              # Spearman rho: exact rank sums vs np.corrcoef
    per_video = eval_fscores(data, users)
    assert bits_equal(np.float64(np.mean(per_video)), np.float64(f))


def test_knapsack_properties_hypothesis():
    """Random small instances (duplicated values force ties): GPU == oracle, selection is feasible,
    ascending, and no single swap-in of an unselected shot that still fits improves the value."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.tuples(st.integers(1, 40), st.sampled_from([0.125, 0.25, 0.5, 0.75, 0.3, 0.7])), min_size=1, max_size=30),
           st.integers(0, 300))
    def check(items, W):
        wt = [w for w, _ in items]
        val = [v for _, v in items]
        got = knapSack(W, wt, val, len(items))
        assert got == c_oracle.knapsack(W, wt, val)
        assert got == sorted(got) and sum(wt[i] for i in got) <= W
        used = sum(wt[i] for i in got)
        for j in range(len(items)):        # values are positive: an optimal selection leaves no room for another shot
            assert j in got or used + wt[j] > W
    check()


def test_empty_and_maximum_sizes():
    # empty inputs
    assert generate_summary([], [], [], []) == []
    assert len(eval_fscores({}, {})) == 0
    # the largest knapsack class: capacity 28 671 frames (n_frames = 191 146 -> N ~ 12 743 sub-sampled frames)
    v = make_video(4000, 12700, n_users=3, with_features=False)
    assert 18944 < int(v.n_frames * 0.15) + 1 <= 28672
    check_against_oracle([v], [make_scores(4000, 12700)], "avg")
    # one frame more than the fp64 DP row can hold in shared memory must fail loudly, not silently
    from vsum_b200 import _cabi
    too_long = make_video(4001, 12800, n_users=1, with_features=False)
    assert int(too_long.n_frames * 0.15) + 1 > 28672
    with pytest.raises(_cabi.VsumError):
        run_batch([too_long], [make_scores(4001, 12800)])


def _scipy_correlations(frame_scores, user_scores):
    """The reference's evaluate_scores (compute_correlation.py:4-15), per user, on fp64 copies so that scipy
    returns its fp64 statistic (with float32 inputs it rounds tau to float32, see _engine.rank_correlations)."""
    from scipy import stats
    taus, rhos = [], []
    pr = stats.rankdata(-frame_scores.astype(np.float64))
    for row in user_scores:
        ur = stats.rankdata(-row.astype(np.float64))
        rhos.append(stats.spearmanr(pr, ur)[0])
        taus.append(stats.kendalltau(pr, ur)[0])
    return np.asarray(taus, np.float64), np.asarray(rhos, np.float64)


def _corr_case(kind, n_steps, n_users, seed):
    from vsum_b200.evaluation.compute_metrics import upsample
    rng = np.random.default_rng(seed)
    n_frames = 15 * (n_steps - 1) + 1 + int(rng.integers(0, 15))
    picks = np.arange(0, n_frames, 15, dtype=np.int32)[:n_steps]
    scores = rng.random(n_steps, dtype=np.float32)
    users = rng.random((n_users, n_frames), dtype=np.float32)
    if kind == "levels":                                   # TVSum-like annotations 1..5 held for 2 s: heavy ties
        users = np.repeat(rng.integers(1, 6, (n_users, n_frames // 30 + 1)), 30, axis=1)[:, :n_frames].astype(np.float32)
    elif kind == "pred_ties":                              # equal scores in different segments, signed zeros
        scores = rng.choice(np.asarray([0.25, 0.5, 0.75, 0.0, -0.0], np.float32), n_steps)
        users[:, ::3] = 0.5
    elif kind == "constant_user":
        users[0] = 2.0                                     # scipy: nan for that user -> nan mean
    elif kind == "head":                                   # frames before the first pick keep np.zeros' 0
        picks = picks + 7
        picks[-1] = min(picks[-1], n_frames - 1)
    elif kind == "perfect":
        users[0] = upsample(scores, n_frames, picks)       # tau = rho = 1 for user 0
        users[1] = -users[0]
    return scores, picks, n_frames, users


@pytest.mark.parametrize("kind,n_steps,n_users", [
    ("random", 40, 3), ("random", 300, 20), ("levels", 300, 20), ("pred_ties", 257, 5), ("constant_user", 64, 4),
    ("head", 90, 3), ("perfect", 120, 3), ("random", 1, 2), ("random", 2, 2), ("levels", 2100, 4), ("random", 8000, 2)])
def test_rank_correlations_match_scipy(kind, n_steps, n_users):
    """vsum_rank_correlation against scipy (the reference's implementation): tau bit-exact per user, rho to 1e-13."""
    from vsum_b200.evaluation import _engine
    from vsum_b200.evaluation.compute_metrics import upsample
    scores, picks, n_frames, users = _corr_case(kind, n_steps, n_users, seed=n_steps * 7 + n_users)
    want_tau, want_rho = _scipy_correlations(upsample(scores, n_frames, picks), users)
    tau, rho, pu_tau, pu_rho = _engine.rank_correlations([scores], [picks], [n_frames], [users], per_user=True)
    assert np.array_equal(np.isnan(pu_tau), np.isnan(want_tau)) and np.array_equal(np.isnan(pu_rho), np.isnan(want_rho))
    ok = ~np.isnan(want_tau)
    assert (pu_tau[ok].view(np.int64) == want_tau[ok].view(np.int64)).all(), (pu_tau, want_tau)
    ok = ~np.isnan(want_rho)
    np.testing.assert_allclose(pu_rho[ok], want_rho[ok], rtol=0, atol=1e-13)
    want32 = list(want_tau.astype(np.float32))            # float32 user scores: scipy rounds each tau to float32
    assert bits_equal(np.float64(tau[0]), np.float64(sum(want32) / len(want32)))          # compute_correlation.py:15
    t64, _ = _engine.rank_correlations([scores], [picks], [n_frames], [users.astype(np.float64)])
    assert bits_equal(np.float64(t64[0]), np.float64(sum(list(want_tau)) / len(want_tau)))
    if kind == "perfect":
        assert pu_tau[0] == 1.0 and pu_tau[1] == -1.0 and abs(pu_rho[0] - 1.0) < 1e-15


def test_rank_correlations_batched_and_dropin():
    """Several videos of different lengths in one call == one call per video; the drop-in evaluate_scores
    (frame-level prediction in, compute_correlation.py:4) gives the same numbers."""
    from vsum_b200.evaluation import _engine
    from vsum_b200.evaluation.compute_correlation import evaluate_scores
    from vsum_b200.evaluation.compute_metrics import upsample
    cases = [_corr_case(k, n, u, seed=5 + i) for i, (k, n, u) in enumerate([("random", 33, 2), ("levels", 700, 20), ("pred_ties", 129, 7),
                                                                             ("random", 1500, 1)])]
    tau, rho = _engine.rank_correlations([c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], [c[3] for c in cases])
    for i, (scores, picks, n_frames, users) in enumerate(cases):
        t1, r1 = _engine.rank_correlations([scores], [picks], [n_frames], [users])
        assert bits_equal(np.float64(tau[i]), np.float64(t1[0])) and bits_equal(np.float64(rho[i]), np.float64(r1[0]))
        t2, r2 = evaluate_scores(upsample(scores, n_frames, picks), users)
        assert bits_equal(np.float64(t2), np.float64(tau[i])) and bits_equal(np.float64(r2), np.float64(rho[i]))
        wt, wr = _scipy_correlations(upsample(scores, n_frames, picks), users)
        wt32 = list(wt.astype(np.float32))
        assert bits_equal(np.float64(tau[i]), np.float64(sum(wt32) / len(wt32))) and abs(rho[i] - sum(wr) / len(wr)) < 1e-13


def test_summary_frames_match_mask():
    """vsum_summary_frames == [i for i, is_in in enumerate(summary) if is_in == 1] (generate_summary_image.py:74-76)."""
    from vsum_b200.evaluation.generate_summary import summary_frames
    vids = [make_video(900 + i, n, n_users=2, with_features=False) for i, n in enumerate((5, 64, 300, 1333, 4000))]
    scores = [make_scores(900 + i, v.n_steps) for i, v in enumerate(vids)]
    args = ([v.change_points for v in vids], scores, [v.n_frames for v in vids], [v.picks for v in vids])
    masks = generate_summary(*args)
    frames = summary_frames(*args)
    for m, f in zip(masks, frames):
        assert f == [i for i, is_in in enumerate(m) if is_in == 1]
    assert any(len(f) > 0 for f in frames)
