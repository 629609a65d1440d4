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


def summary_frames(all_shot_bound, all_scores, all_nframes, all_positions):
    """Frame numbers of every video's summary, ascending -- the lists `generate_summary_image.get_summary`
    (src/generate_summary_image.py:54-80) stores in summary.json -- without building the int8 masks:
    `vsum_summary_frames` expands the selected shots on the GPU."""
    B = len(all_scores)
    if B == 0:
        return []
    hb = _engine.HostEvalBatch.build(all_shot_bound, all_nframes, all_positions)
    db = _engine.DeviceEvalBatch(hb)
    dev = db.device
    scores = [np.ascontiguousarray(np.asarray(s), dtype=np.float32).reshape(-1) for s in all_scores]
    cu_steps = torch.from_numpy(_engine._cu([len(s) for s in scores]).astype(np.int32)).to(dev)
    d_scores = torch.from_numpy(np.concatenate(scores)).to(dev)
    out = _engine.summarize(db, d_scores, cu_steps, want_f=False)
    caps = [_engine.capacity_of(int(c[-1, 1])) if len(c) else 0
            for c in (np.asarray(c).reshape(-1, 2) for c in all_shot_bound)]
    off = _engine._cu(caps)
    with torch.cuda.device(dev):
        d_off = torch.from_numpy(off).to(dev)
        frames = torch.empty(max(int(off[-1]), 1), dtype=torch.int32, device=dev)
        counts = torch.empty(B, dtype=torch.int32, device=dev)
        _engine._cabi.check(_engine._cabi.load().vsum_summary_frames(
            out["selected"].data_ptr(), db.cps.data_ptr(), db.cu_shots.data_ptr(), d_off.data_ptr(), B, frames.data_ptr(),
            counts.data_ptr(), torch.cuda.current_stream(dev).cuda_stream), "vsum_summary_frames")
        frames, counts = frames.cpu().numpy(), counts.cpu().numpy()
    return [frames[off[v]:off[v] + counts[v]].tolist() for v in range(B)]


def get_summary(model, data_loader):
    """`generate_summary_image.get_summary` (src/generate_summary_image.py:54-80): scores every video of the
    loader with the model and returns {"video_i": [frame numbers]} ready for json.dump."""
    model.eval()
    scores, users = [], []
    with torch.no_grad():
        for feature, _target, user in data_loader:
            pred, _ = model(feature.to(next(model.parameters()).device))
            scores.append(torch.sigmoid(pred.view(1, -1)).squeeze(0).cpu().numpy())
            users.append(user)
    lists = summary_frames([u.change_points for u in users], scores, [u.n_frames for u in users], [u.picks for u in users])
    return {"video_%d" % i: frames for i, frames in enumerate(lists)}
