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

from . import _cabi
from .evaluation import _engine
from .model.simnet import SimNet
from .synthetic import SyntheticVideo


@dataclass
class HostBatch:
    features: torch.Tensor            # float32 (or bfloat16, see data.write_pack) [T,1024] (pinned when pin=True)
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

    def __init__(self, hb: HostBatch, device=None, features: Optional[torch.Tensor] = None, pin_meta: bool = False):
        self.host = hb
        self.meta = _engine.DeviceEvalBatch(hb.meta, device, pin=pin_meta)
        dev = self.meta.device
        self.features = features if features is not None else hb.features.to(dev, non_blocking=True)
        self._cu_host = torch.from_numpy(hb.cu_steps)
        if pin_meta:
            self._cu_host = self._cu_host.pin_memory()
        self.cu_steps = self._cu_host.to(dev, non_blocking=True)
        self.h2d_bytes = self.meta.h2d_bytes + hb.features.numel() * hb.features.element_size() + hb.cu_steps.nbytes

    def refill(self) -> None:
        """Copy features + metadata of the host batch into the existing device buffers again."""
        self.features.copy_(self.host.features, non_blocking=True)
        self.cu_steps.copy_(self._cu_host, non_blocking=True)
        self.meta.refill()


class _Slot:
    def __init__(self):
        self.buf = None         # fp32 [>=T,1] score buffer owned by the slot
        self.scores = None      # its [:T] view for the current batch
        self.done = None        # event: evaluation of the batch that last used this slot finished
        self.f_host = None      # pinned fp64 [B] landing buffer (end-to-end path)
        self.copied = None      # event: H2D of this slot finished


class Summarizer:
    """Runs the path for batches of packed videos.  Two execution styles:

    * `run_device` / `run_host`: one batch, everything on the current stream (used by the parity tests);
    * `submit_device` / `submit_host` + `drain`: software pipeline over consecutive batches -- the
      scorer of batch k+1 runs on the main stream while shot pooling / knapsack / F-score of batch k
      run on a side stream (the knapsack of the longest video is a serial ~ms tail on ONE SM), and,
      for host batches, the H2D copy of batch k+1 runs on a copy stream.
    """

    def __init__(self, model: SimNet, eval_method: str = "avg", eval_sms: int = 0):
        self.model = model
        self.eval_method = eval_method
        # optional SM partition for pipelined mode (0 = off, the measured optimum on the bench workload):
        # pooling / knapsack / F-score keep at most `eval_sms` SMs busy and the scorer's
        # persistent GEMMs leave as many free, so neither stream waits for an SM the other holds
        self.eval_sms = eval_sms
        self._side = None
        self._copy = None
        self._slots = {}

    @torch.no_grad()
    def run_device(self, db: DeviceBatch, want_intermediates: bool = False):
        """Inputs already in HBM.  Returns the per-video F tensor (fp64, packed order) or the
        full dict of intermediates."""
        _cabi.check(_cabi.load().vsum_set_sm_partition(0, 0), "vsum_set_sm_partition")   # one stream: no partition
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

    # ------------------------------------------------------------------ pipelined execution
    def _streams(self, dev):
        if self._side is None:
            self._side = torch.cuda.Stream(dev)
            self._copy = torch.cuda.Stream(dev)
        return torch.cuda.current_stream(dev), self._side, self._copy

    def _slot(self, slot: int, db: DeviceBatch) -> _Slot:
        """Per-slot score buffer, grow-only (batches of different sizes share it through a [:T] view)."""
        st = self._slots.setdefault(slot, _Slot())
        T = db.features.shape[0]
        if st.buf is None or st.buf.shape[0] < T or st.buf.device != db.features.device:
            if st.buf is not None and self._side is not None:
                st.buf.record_stream(self._side)              # the evaluation stream may still read the outgrown buffer
            st.buf = torch.empty((int(T * 1.25) + 1, 1), dtype=torch.float32, device=db.features.device)
        st.scores = st.buf[:T]
        return st

    @torch.no_grad()
    def submit_device(self, db: DeviceBatch, slot: int, to_host: bool = False) -> torch.Tensor:
        """Queue one resident batch (a `DeviceBatch`, or an `EvalBatch` from `data.PackedEvalLoader`, whose device slot
        is handed back to the loader when the evaluation has consumed it).  Returns the per-video F tensor, valid after
        `drain()` (or after waiting on the slot); with `to_host` the F-scores are also copied into a pinned fp64 buffer
        on the side stream and THAT tensor is returned.  Use two slots alternately."""
        dev = db.features.device
        main, side, _ = self._streams(dev)
        st = self._slot(slot, db)
        _cabi.check(_cabi.load().vsum_set_sm_partition(self.eval_sms, self.eval_sms), "vsum_set_sm_partition")
        if st.done is not None:
            main.wait_event(st.done)                      # the score buffer of this slot is free again
        if st.copied is not None:
            main.wait_event(st.copied)
        self.model.forward_packed(db.features, db.cu_steps, db.host.seqlens, apply_sigmoid=True,
                                  want_feats=False, scores_out=st.scores)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ready)
            f = _engine.summarize(db.meta, st.scores.view(-1), db.cu_steps, self.eval_method)["f"]
            if to_host:
                if st.f_host is None or st.f_host.shape[0] < f.shape[0]:
                    st.f_host = torch.empty(f.shape[0], dtype=torch.float64, pin_memory=True)
                out = st.f_host[:f.shape[0]]
                out.copy_(f, non_blocking=True)
                f = out
            st.done = torch.cuda.Event()
            st.done.record(side)
        if hasattr(db, "release"):
            db.release(st.done)
        return f

    @torch.no_grad()
    def submit_host(self, db: DeviceBatch, slot: int) -> torch.Tensor:
        """Queue one HOST batch through device buffers `db` (built with pin_meta=True): H2D on the
        copy stream, compute as in submit_device, D2H of the F-scores into a pinned buffer on the
        side stream.  Returns the pinned fp64 [B] tensor (packed order), valid after `drain()`."""
        dev = db.features.device
        main, side, copy = self._streams(dev)
        st = self._slot(slot, db)
        with torch.cuda.stream(copy):
            if st.done is not None:
                copy.wait_event(st.done)                  # previous user of these device buffers finished
            db.refill()
            st.copied = torch.cuda.Event()
            st.copied.record(copy)
        f = self.submit_device(db, slot)
        if st.f_host is None or st.f_host.shape[0] < f.shape[0]:
            st.f_host = torch.empty(f.shape[0], dtype=torch.float64, pin_memory=True)
        out = st.f_host[:f.shape[0]]
        with torch.cuda.stream(side):
            out.copy_(f, non_blocking=True)
            st.done = torch.cuda.Event()
            st.done.record(side)
        return out

    def drain(self, dev=None) -> None:
        """Make the current stream wait for everything queued on the side / copy streams."""
        if self._side is not None:
            main = torch.cuda.current_stream(dev)
            main.wait_stream(self._side)
            main.wait_stream(self._copy)
