"""Batched summarise-and-evaluate pass: the body of the reference's `train.py:val_step`
(lines 134-152) for MANY videos at once instead of one video per Python iteration.

    features -> SimNet scorer (sigmoid fused) -> shot pooling -> knapsack -> mask -> F-score

Videos are packed without padding (longest first, so the attention tile list starts with the
heavy work) and every stage is one or a few sm_100a kernel launches over the whole batch.
`Summarizer.run_host` is the end-to-end call the bench times: pinned host buffers in, per-video
F-scores out, with both copies inside the call.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from .evaluation import _engine
from .model.simnet import SimNet
from .synthetic import SyntheticVideo


@dataclass
class HostBatch:
    features: torch.Tensor            # float32 [T,1024] (pinned when pin=True)
    seqlens: List[int]                # per packed video
    cu_steps: np.ndarray              # int32[B+1]
    meta: _engine.HostEvalBatch
    order: np.ndarray                 # packed position -> index in the caller's list
    names: List[str]

    @property
    def n_videos(self) -> int:
        return len(self.seqlens)

    @property
    def n_steps(self) -> int:
        return int(self.cu_steps[-1])


def pack_videos(videos: Sequence[SyntheticVideo], pin: bool = True, with_features: bool = True) -> HostBatch:
    order = np.argsort([-v.n_steps for v in videos], kind="stable")
    vs = [videos[i] for i in order]
    seqlens = [v.n_steps for v in vs]
    cu = _engine._cu(seqlens).astype(np.int32)
    feats = torch.empty((int(cu[-1]), 1024), dtype=torch.float32, pin_memory=pin and torch.cuda.is_available())
    if with_features:
        fn = feats.numpy()
        for v, a, b in zip(vs, cu[:-1], cu[1:]):
            fn[a:b] = v.features
    meta = _engine.HostEvalBatch.build([v.change_points for v in vs], [v.n_frames for v in vs],
                                       [v.picks for v in vs], [v.user_summary for v in vs])
    return HostBatch(feats, seqlens, cu, meta, order, [v.name for v in vs])


class DeviceBatch:
    """A HostBatch resident in HBM."""

    def __init__(self, hb: HostBatch, device=None, features: Optional[torch.Tensor] = None):
        self.host = hb
        self.meta = _engine.DeviceEvalBatch(hb.meta, device, pin=False)
        dev = self.meta.device
        self.features = features if features is not None else hb.features.to(dev, non_blocking=True)
        self.cu_steps = torch.from_numpy(hb.cu_steps).to(dev, non_blocking=True)
        self.h2d_bytes = self.meta.h2d_bytes + hb.features.numel() * 4 + hb.cu_steps.nbytes


class Summarizer:
    def __init__(self, model: SimNet, eval_method: str = "avg"):
        self.model = model
        self.eval_method = eval_method

    @torch.no_grad()
    def run_device(self, db: DeviceBatch, want_intermediates: bool = False):
        """Inputs already in HBM.  Returns the per-video F tensor (fp64, packed order) or the
        full dict of intermediates."""
        scores, _ = self.model.forward_packed(db.features, db.cu_steps, db.host.seqlens,
                                              apply_sigmoid=True, want_feats=False)   # train.py:143-144
        out = _engine.summarize(db.meta, scores.view(-1), db.cu_steps, self.eval_method)
        if want_intermediates:
            out["scores"] = scores.view(-1)
            return out
        return out["f"]

    @torch.no_grad()
    def run_host(self, hb: HostBatch, device=None) -> np.ndarray:
        """End to end from host buffers: H2D of features + metadata, the whole path, D2H of the
        per-video F-scores (returned in the caller's original video order)."""
        db = DeviceBatch(hb, device)
        f_packed = self.run_device(db).cpu().numpy()
        f = np.empty_like(f_packed)
        f[hb.order] = f_packed
        return f
