"""CPU, build container only: live differential test of the oracle against the reference
imported from /root/reference (skipped where the reference is not mounted, e.g. the GPU box)."""
import importlib
import os
import sys
import warnings

import numpy as np
import pytest

from conftest import REFERENCE, bits_equal
from oracle import c_oracle, ref_port
from vsum_b200.synthetic import make_scores, make_video

pytestmark = pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted")


@pytest.fixture(scope="module")
def ref_eval():
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("evaluation",)}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        mods = dict(gs=importlib.import_module("evaluation.generate_summary"),
                    ks=importlib.import_module("evaluation.knapsack_implementation"),
                    em=importlib.import_module("evaluation.evaluation_metrics"),
                    cm=importlib.import_module("evaluation.compute_metrics"))
        assert mods["gs"].__file__.startswith(REFERENCE)
        yield mods
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k.split(".")[0] == "evaluation"]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_videos_against_live_reference(ref_eval):
    rng = np.random.default_rng(11)
    for v_id in range(300, 340):
        n = int(rng.integers(1, 700))
        users = int(rng.integers(1, 21))
        v = make_video(v_id, n, n_users=users, with_features=False)
        sc = make_scores(v_id, n)
        ref_summary = ref_eval["gs"].generate_summary([v.change_points], [sc], [np.array(v.n_frames)], [v.picks])[0]
        c = c_oracle.video(sc, v.picks, v.n_frames, v.change_points, v.user_summary, "avg")
        assert bits_equal(ref_summary, c["summary"])
        py_summary, means, *_ = ref_port.summarize_video(v.change_points, sc, v.n_frames, v.picks)
        assert bits_equal(ref_summary, py_summary) and bits_equal(means, c["val"])
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for method in ("avg", "max"):
                want = np.float64(ref_eval["em"].evaluate_summary(ref_summary, v.user_summary, method))
                assert bits_equal(want, np.float64(c_oracle.fscore(ref_summary, v.user_summary, method)[0]))
                assert bits_equal(want, np.float64(ref_port.fscore_video(ref_summary, v.user_summary, method)))
                if v.n_steps <= 400:
                    assert bits_equal(want, np.float64(ref_port.fscore_video(ref_summary, v.user_summary, method, builtin_sums=True)))


def test_compact_input_forms_are_lossless_for_the_reference(ref_eval):
    """The packed dataset may hold the 0/1 user summaries as uint8: the reference's own evaluate_summary returns the same
    bits for either dtype (it copies the row into an int array, evaluation_metrics.py:13-19), which is what makes
    `write_pack(user_summary_u8=True)` a lossless storage choice rather than a change of inputs."""
    for v_id, n, users in ((350, 90, 3), (351, 400, 20), (352, 7, 1)):
        v = make_video(v_id, n, n_users=users, with_features=False)
        summary = ref_eval["gs"].generate_summary([v.change_points], [make_scores(v_id, n)], [np.array(v.n_frames)], [v.picks])[0]
        for method in ("avg", "max"):
            f32 = np.float64(ref_eval["em"].evaluate_summary(summary, v.user_summary, method))
            u8 = np.float64(ref_eval["em"].evaluate_summary(summary, v.user_summary.astype(np.uint8), method))
            assert bits_equal(f32, u8)


def test_random_knapsacks_against_live_reference(ref_eval):
    rng = np.random.default_rng(12)
    for _ in range(200):
        n = int(rng.integers(1, 25))
        wt = rng.integers(1, 30, n)
        # duplicated values on purpose: exercises the tie rule of line 26
        val = [float(x) for x in rng.choice(rng.random(max(2, n // 2)), n)]
        W = int(rng.integers(0, 80))
        want = ref_eval["ks"].knapSack(W, list(wt), val, n)
        assert c_oracle.knapsack(W, wt, val) == want
        assert ref_port.knapsack_select(W, list(wt), val, n) == want


def test_upsample_host_helper(ref_eval):
    from vsum_b200.evaluation.compute_metrics import upsample
    rng = np.random.default_rng(13)
    for v_id in range(20):
        n = int(rng.integers(1, 400))
        v = make_video(v_id, n, with_features=False)
        sc = make_scores(v_id, n)
        assert bits_equal(ref_eval["cm"].upsample(sc, v.n_frames, v.picks), upsample(sc, v.n_frames, v.picks))
        assert bits_equal(ref_eval["cm"].upsample(sc, v.n_frames, v.picks), ref_port.upsample_scores(sc, v.n_frames, v.picks))


def test_kts_port_against_live_reference():
    """oracle/kts_ref.py == the reference's KTS (loaded as a stand-alone package: `data` itself needs h5py)."""
    import contextlib, importlib.util, io
    from oracle import kts_ref
    d = os.path.join(REFERENCE, "data/preprocess/segmentations/kts")
    spec = importlib.util.spec_from_file_location("ref_kts_live", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_kts_live"] = ref
    spec.loader.exec_module(ref)
    for seed, (n, dim, ncp, kw) in enumerate([(70, 16, 12, {}), (45, 8, 6, dict(lmin=2, lmax=15)), (20, 4, 19, {}), (90, 32, 1, {})]):
        rng = np.random.default_rng(100 + seed)
        x = rng.random((n, dim), dtype=np.float32)
        x[n // 3:] += 0.7
        K = np.dot(x, x.T)
        with contextlib.redirect_stdout(io.StringIO()):
            want = ref.kts_segmentation(K, ncp, 0.8, **kw)
        got = kts_ref.kts_segmentation(K, ncp, 0.8, **kw)
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1].view(np.int64), want[1].view(np.int64))
