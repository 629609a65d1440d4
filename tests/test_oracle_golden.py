"""CPU: the oracle (pure-Python port, C restatement, torch fp32 scorer restatement) against the
golden fixtures that the imported reference produced (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import bits_equal
from oracle import c_oracle, ref_port, scorer_ref
from vsum_b200.synthetic import make_scores, make_video


def _cases(golden):
    return [tuple(int(x) for x in row) for row in golden["cases"]]


def test_knapsack_known_answers(eval_golden):
    # the reference's only shipped known answer: knapsack_implementation.py:35-41
    assert list(eval_golden["knap_driver"]) == [0, 1, 2, 3, 4]
    for impl in (ref_port.knapsack_select, lambda W, wt, val, n: c_oracle.knapsack(W, wt, val)):
        assert impl(7, [2, 2, 1, 1, 1, 2], [4, 4, 2, 2, 2, 4], 6) == [0, 1, 2, 3, 4]
        assert impl(2, [2, 1, 1], [2, 1, 1], 3) == list(eval_golden["knap_tie_a"]) == [0]
        assert impl(2, [1, 1, 2], [1, 1, 2], 3) == list(eval_golden["knap_tie_b"]) == [0, 1]


def test_c_oracle_matches_golden(eval_golden):
    for vid, n, users in _cases(eval_golden):
        v = make_video(vid, n, n_users=users, with_features=False)
        sc = make_scores(vid, n)
        for k, method in enumerate(("avg", "max")):
            r = c_oracle.video(sc, v.picks, v.n_frames, v.change_points, v.user_summary, method)
            want_summary = np.unpackbits(eval_golden[f"summary_{vid}"])[:len(r["summary"])].astype(np.int8)
            assert bits_equal(r["summary"], want_summary), (vid, n)
            assert bits_equal(r["val"], eval_golden[f"means_{vid}"]), (vid, n)
            assert bits_equal(np.float64(r["f"]), np.float64(eval_golden[f"f_{vid}"][k])), (vid, n, method)


def test_python_port_matches_golden(eval_golden):
    for vid, n, users in _cases(eval_golden):
        if n > 1300:          # the pure-Python DP is the reference's own speed: keep the CPU suite short
            continue
        v = make_video(vid, n, n_users=users, with_features=False)
        summary, means, lengths, cap, _ = ref_port.summarize_video(v.change_points, make_scores(vid, n), v.n_frames, v.picks)
        want = np.unpackbits(eval_golden[f"summary_{vid}"])[:len(summary)].astype(np.int8)
        assert bits_equal(summary, want)
        assert bits_equal(means, eval_golden[f"means_{vid}"])
        for k, method in enumerate(("avg", "max")):
            f = ref_port.fscore_video(summary, v.user_summary, method)
            assert bits_equal(np.float64(f), np.float64(eval_golden[f"f_{vid}"][k]))
            if n <= 400:      # the timing arm's form: builtin sum() counts like evaluation_metrics.py:23-24 -- same bits, ~100x the time
                assert bits_equal(np.float64(ref_port.fscore_video(summary, v.user_summary, method, builtin_sums=True)), np.float64(f))
        if n <= 400:          # ... and the multiprocessing worker of bench.py's `fair` CPU figure
            got = ref_port.pool_eval_video((v.change_points, make_scores(vid, n), v.n_frames, v.picks, v.user_summary, "avg"))
            assert bits_equal(np.float64(got), np.float64(eval_golden[f"f_{vid}"][0]))


def test_pairwise_sum_matches_numpy():
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [127, 128, 129, 255, 256, 257, 1000, 4099, 123457]:
        a = (rng.random(n, dtype=np.float32) * np.float32(rng.choice([1.0, 100.0, 1e-3]))).astype(np.float32)
        got = np.float32(np.float32(0.0) + c_oracle.pairwise_sum_f32(a))
        assert bits_equal(got, np.add.reduce(a)), n
        if n < 600:
            assert bits_equal(ref_port.pairwise_sum_f32(a, 0, n), c_oracle.pairwise_sum_f32(a)), n


def test_nan_when_nothing_fits():
    # a single shot longer than 15 % of the video: empty summary -> 0/0 -> NaN
    cps = np.array([[0, 99]], dtype=np.int32)
    us = np.ones((2, 100), dtype=np.float32)
    r = c_oracle.video(np.full(7, 0.5, np.float32), np.arange(0, 100, 15, dtype=np.int32), 100, cps, us)
    assert r["summary"].sum() == 0 and np.isnan(r["f"])


@pytest.mark.parametrize("vid,n", [(100, 1), (101, 5), (102, 130), (103, 300)])
def test_scorer_restatement_matches_reference(scorer_golden, seeded_model_kwargs, vid, n):
    from vsum_b200.model import SimNet
    torch.manual_seed(1234)
    import random; random.seed(1234); np.random.seed(1234)
    sd = SimNet(**seeded_model_kwargs).state_dict()
    x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0)
    logits, feats = scorer_ref.scorer_forward(sd, x, num_heads=4)
    np.testing.assert_allclose(logits.view(-1).numpy(), scorer_golden[f"logits_{vid}"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(feats[0, :4].numpy(), scorer_golden[f"feats_head_{vid}"], rtol=1e-5, atol=1e-5)


def test_masked_mse_restatement_matches_reference_loss(scorer_golden):
    """oracle/scorer_ref.masked_mse against the loss the reference's `mse_with_mask_loss` (utils.py:45-56) returned on its own
    padded-batch logits (tests/golden/make_golden.py: padded_loss)."""
    lens = (300, 180, 77)
    logits = torch.full((3, 300, 1), -3.0)
    tgt = torch.full((3, 300), 1000.0)
    mask = torch.ones((3, 300), dtype=torch.bool)
    for b, n in enumerate(lens):
        logits[b, :n, 0] = torch.from_numpy(scorer_golden[f"padded_logits_{b}"])
        tgt[b, :n] = torch.from_numpy(make_video(110 + b, n).gtscore)
        mask[b, :n] = False
    np.testing.assert_allclose(scorer_ref.masked_mse(logits, tgt, mask).item(), float(scorer_golden["padded_loss"]), rtol=1e-6)


@pytest.mark.parametrize("vid,n", [(107, 4096)])
def test_scorer_restatement_matches_reference_at_full_length(scorer_long_golden, seeded_model_kwargs, vid, n):
    """The torch restatement (with the extended sinusoid table) against the reference at N = 4096."""
    from vsum_b200.model import SimNet
    torch.manual_seed(1234)
    sd = SimNet(**seeded_model_kwargs).state_dict()
    x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0)
    logits, _ = scorer_ref.scorer_forward(sd, x, num_heads=4)
    np.testing.assert_allclose(logits.view(-1).numpy(), scorer_long_golden[f"logits_{vid}"], rtol=1e-5, atol=1e-5)


def test_bf16_feature_rounding_stays_far_inside_the_score_tolerance(seeded_model_kwargs):
    """What `write_pack(features_bf16=True)` does to the inputs, seen through the fp32 restatement of the reference
    scorer: rounding the features to bfloat16 once moves the importance scores by a few 1e-4 relative, against the 1e-2
    the bf16 path is allowed (BASELINE.json) -- the input rounding is a small part of that budget."""
    from vsum_b200.model import SimNet
    torch.manual_seed(1234)
    sd = SimNet(**seeded_model_kwargs).state_dict()
    worst = 0.0
    for vid, n in ((120, 64), (121, 300), (122, 700)):
        x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0)
        a = torch.sigmoid(scorer_ref.scorer_forward(sd, x, num_heads=4)[0]).view(-1)
        b = torch.sigmoid(scorer_ref.scorer_forward(sd, x.bfloat16().float(), num_heads=4)[0]).view(-1)
        worst = max(worst, float(((a - b).abs() / a).max()))
    assert worst < 2e-3, worst


def test_kts_restatement_matches_reference_golden():
    """oracle/kts_ref.py against the reference's own kts_segmentation / cpd_nonlin outputs (tests/golden/kts_golden.npz,
    written by make_golden.py from src/data/preprocess/segmentations/kts): bit-exact costs, same change points."""
    import os
    from conftest import GOLDEN
    from oracle import kts_ref
    g = np.load(os.path.join(GOLDEN, "kts_golden.npz"))
    for seed, n, dim, ncp, lmin, lmax in g["cases"]:
        K = g[f"K_{seed}"]
        cps, costs = kts_ref.kts_segmentation(K, int(ncp), 1.0, lmin=int(lmin), lmax=int(lmax))
        assert np.array_equal(cps, g[f"cps_{seed}"])
        assert np.array_equal(costs.view(np.int64), g[f"costs_{seed}"].view(np.int64))
        cps2, scores = kts_ref.cpd_nonlin(K, int(ncp), lmin=int(lmin), lmax=int(lmax))
        assert np.array_equal(cps2, g[f"cpsfixed_{seed}"])
        assert np.array_equal(scores.view(np.int64), g[f"scores_{seed}"].view(np.int64))
