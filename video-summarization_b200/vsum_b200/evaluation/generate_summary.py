"""`generate_summary` with the reference signature (`src/evaluation/generate_summary.py:6-57`),
computed on the GPU: shot pooling, knapsack and the int8 keyshot mask are the sm_100a kernels
behind `vsum_shot_mean`, `vsum_knapsack` and `vsum_summary_fscore`."""
from __future__ import annotations

import numpy as np
import torch

from . import _engine


def generate_summary(all_shot_bound, all_scores, all_nframes, all_positions):
    """Lists (one entry per video) of change points int[S,2], scores float32[N], n_frames and
    picks.  Returns a list of `np.int8` masks of length `last_shot_end + 1`."""
    B = len(all_scores)
    if B == 0:
        return []
    hb = _engine.HostEvalBatch.build(all_shot_bound, all_nframes, all_positions)
    db = _engine.DeviceEvalBatch(hb)
    scores = [np.ascontiguousarray(np.asarray(s), dtype=np.float32).reshape(-1) for s in all_scores]
    cu_steps = torch.from_numpy(_engine._cu([len(s) for s in scores]).astype(np.int32)).to(db.device)
    d_scores = torch.from_numpy(np.concatenate(scores)).to(db.device)
    out = _engine.summarize(db, d_scores, cu_steps, want_f=False)
    flat = out["summary"].cpu().numpy()
    off = hb.sum_offsets
    return [flat[off[v]:off[v + 1]].copy() for v in range(B)]
