"""Rank correlations against each user's frame scores with the reference signature
(`src/evaluation/compute_correlation.py:4-15`), computed by `vsum_rank_correlation` on the GPU
(SURVEY.md section 8(f) rank 1).  Kendall's tau is bit-exact with scipy; Spearman's rho agrees to ~1e-15."""
import numpy as np

from . import _engine


def evaluate_scores(predicted_summary, user_scores):
    """predicted_summary: float32[n_frames] (upsampled frame scores); user_scores [U, n_frames].
    Returns (mean Kendall tau-b, mean Spearman rho) over the users.
    Limit: the prediction is handed to the kernels as runs of equal values (what `upsample` produces from per-pick scores,
    compute_metrics.py:19-39); a prediction with more than 8191 distinct runs (e.g. a dense per-frame score vector on a
    video longer than ~8k frames) raises VSUM_EUNSUPPORTED where the reference (scipy) would accept it."""
    pred = np.ascontiguousarray(np.asarray(predicted_summary), dtype=np.float32).reshape(-1)
    n = len(pred)
    # the kernels take the piecewise-constant form: run starts are the picks, run values the scores
    starts = np.concatenate([[0], np.flatnonzero(pred[1:] != pred[:-1]) + 1]).astype(np.int32) if n else np.zeros(0, np.int32)
    tau, rho = _engine.rank_correlations([pred[starts]], [starts], [n], [np.asarray(user_scores)])
    return tau[0], rho[0]
