"""`eval_metrics` / `upsample` with the reference signatures
(`src/evaluation/compute_metrics.py:19-39, 42-92`).

`eval_metrics(data, user_dict)` takes the dict of per-video float32 score arrays that
`train.py:val_step` builds and the matching `UserSummaries` records, runs shot pooling, knapsack,
mask and F-score for all videos in ONE batched pass on the GPU and returns
`(mean F, mean Kendall tau, mean Spearman rho)`.  The mean over videos is `np.mean` on the host
in dict insertion order, exactly like line 92, so it does not depend on how videos were batched
or sharded.
"""
from __future__ import annotations

import logging

import numpy as np
import torch

from . import _engine
from .compute_correlation import evaluate_scores


def upsample(scores, n_frames, positions):
    """Repeat each sub-sampled score up to the next pick (host helper with the reference signature;
    the GPU kernels never materialise this array -- see vsum_shot_mean / vsum_rank_correlation)."""
    n_frames = int(n_frames)
    pos = np.asarray(positions)
    if pos.dtype != int:
        pos = pos.astype(np.int32)
    if pos[-1] != n_frames:
        pos = np.concatenate([pos, [n_frames]])
    out = np.zeros(n_frames, dtype=np.float32)
    seg = np.diff(np.clip(pos, 0, n_frames))
    vals = np.zeros(len(seg), dtype=np.float32)
    k = min(len(seg), len(scores))
    vals[:k] = np.asarray(scores, dtype=np.float32)[:k]
    lo = int(np.clip(pos[0], 0, n_frames))
    body = np.repeat(vals, np.maximum(seg, 0))
    out[lo:lo + len(body)] = body[:n_frames - lo]
    return out


def eval_fscores(data: dict, user_dict: dict, eval_method: str = "avg", device=None) -> np.ndarray:
    """Per-video F-scores (fp64) in dict order -- the GPU part of `eval_metrics`."""
    keys = list(data.keys())
    if not keys:
        return np.zeros(0, dtype=np.float64)
    users = [user_dict[k] for k in keys]
    hb = _engine.HostEvalBatch.build([u.change_points for u in users], [u.n_frames for u in users],
                                     [u.picks for u in users], [u.user_summary for u in users])
    db = _engine.DeviceEvalBatch(hb, device)
    scores = [np.ascontiguousarray(np.asarray(data[k]), dtype=np.float32).reshape(-1) for k in keys]
    cu_steps = torch.from_numpy(_engine._cu([len(s) for s in scores]).astype(np.int32)).to(db.device)
    d_scores = torch.from_numpy(np.concatenate(scores)).to(db.device)
    return _engine.summarize(db, d_scores, cu_steps, eval_method)["f"].cpu().numpy()


def eval_metrics(data, user_dict, with_correlation: bool = True):
    eval_method = 'avg'                      # the reference hard-codes it (compute_metrics.py:43)
    keys = list(data.keys())
    f_scores = eval_fscores(data, user_dict, eval_method)
    taus, rhos = [], []
    if with_correlation and keys:               # all videos in one batched GPU pass (compute_metrics.py:80-85)
        users = [user_dict[k] for k in keys]
        taus, rhos = _engine.rank_correlations([data[k] for k in keys], [u.picks for u in users],
                                               [u.n_frames for u in users], [u.user_scores for u in users])
    mean_f = np.mean(f_scores)
    mean_tau = np.mean(taus) if len(taus) else float("nan")
    mean_rho = np.mean(rhos) if len(rhos) else float("nan")
    logging.info(f" [f_score: {mean_f:.4f}, kenadall_tau: {mean_tau:.4f}, spearsman_r: {mean_rho:.4f}]")
    return mean_f, mean_tau, mean_rho
