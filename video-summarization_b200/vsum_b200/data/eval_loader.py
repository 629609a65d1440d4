"""Evaluation-side loader: batches of MANY validation videos, DMA'd straight out of a page-locked pack.

Replaces the reference's `val_loader` (a batch-size-1 DataLoader over `TSDataset(split="val")`,
`src/train.py:68-72`) together with the per-video h5 reads (`src/data/dataset.py:89-103,127-135`) feeding
`val_step` (`src/train.py:139-148`).  A batch is built in two parts:

* the small per-video arrays (picks, change points, offsets, knapsack order and launch classes) are gathered by
  the native `vsum_pack_eval_collate` into one pinned blob and cross PCIe with ONE copy;
* the two big arrays -- features and user summaries -- are never gathered on the host: `vsum_pack_h2d` issues one
  `cudaMemcpyAsync` per video from the page-locked dataset (`PackedDataset(resident="pinned")`) into the packed
  device rows.  The copy engine does the collate.

One background thread prepares batch k+1 .. k+slots-1 (blob + copies on the loader's copy stream) while the
consumer computes on batch k; device landing zones are a ring of `slots` buffers sized for the largest batch.
"""
from __future__ import annotations

import ctypes as C
import queue
import threading
import time
from types import SimpleNamespace
from typing import List, Optional

import numpy as np
import torch

from .. import _cabi
from ..evaluation import _engine
from .packed import PackedDataset

_I32, _I64 = np.int32, np.int64


class EvalBatch:
    """One evaluation batch in HBM -- the object `Summarizer.submit_device` / `run_device` consume (same attributes as
    `pipeline.DeviceBatch`).  `release()` hands the device slot back to the loader."""

    def __init__(self, slot: "_Slot", lay: _cabi.EvalBatchLayout, ids: List[int], names: List[str]):
        self._slot, self.layout = slot, lay
        B, T = int(lay.B), int(lay.T)
        hv = lambda off, n, dt: np.frombuffer(slot.blob_np, dtype=dt, count=n, offset=int(off))      # host views (pinned blob)
        dv = lambda off, n, dt: slot.blob_dev[int(off):int(off) + n * dt.itemsize].view(dt)         # device views
        ti32, ti64 = torch.int32, torch.int64
        cu_host = hv(lay.off_cu_steps, B + 1, _I32)
        self.video_ids = hv(lay.off_video_ids, B, _I32).copy()           # packed position -> raw pack index
        raw_to_ds = {int(slot.ds.ids[i]): i for i in ids}
        self.ids = [raw_to_ds[int(v)] for v in self.video_ids]           # ... -> dataset index
        self.names = [slot.ds.names[i] for i in self.ids]
        self.host = SimpleNamespace(seqlens=np.diff(cu_host).tolist(), n_videos=B, n_steps=T, names=self.names,
                                    cu_steps=cu_host.copy())
        us_dt = np.uint8 if lay.user_summary_dtype == _cabi.USER_SUMMARY_U8 else np.float32
        has_users = lay.total_users > 0
        hb = _engine.HostEvalBatch(
            B=B, picks=hv(lay.off_picks, T, _I32), cu_picks=hv(lay.off_cu_picks, B + 1, _I32),
            n_frames=hv(lay.off_n_frames, B, _I32), cps=hv(lay.off_cps, 2 * lay.total_shots, _I32).reshape(-1, 2),
            cu_shots=hv(lay.off_cu_shots, B + 1, _I32).copy(), bit_offsets=hv(lay.off_bit_offsets, B + 1, _I64).copy(),
            order=hv(lay.off_order, B, _I32), max_cap=int(lay.max_cap),
            launches=[(int(lay.launch_first[i]), int(lay.launch_count[i]), int(lay.launch_max_cap[i])) for i in range(lay.n_launches)],
            sum_offsets=hv(lay.off_sum_offsets, B + 1, _I64).copy(),
            user_summary=np.zeros(0, us_dt) if has_users else None,       # dtype marker only: the rows live on the device
            us_offsets=hv(lay.off_us_offsets, B + 1, _I64).copy() if has_users else None,
            cu_users=hv(lay.off_cu_users, B + 1, _I32).copy() if has_users else None,
            us_cols=hv(lay.off_us_cols, B, _I32) if has_users else None)
        meta = SimpleNamespace(host=hb, device=slot.device, has_users=has_users,
                               picks=dv(lay.off_picks, T, ti32), cu_picks=dv(lay.off_cu_picks, B + 1, ti32),
                               n_frames=dv(lay.off_n_frames, B, ti32), cps=dv(lay.off_cps, 2 * lay.total_shots, ti32).view(-1, 2),
                               cu_shots=dv(lay.off_cu_shots, B + 1, ti32), bit_offsets=dv(lay.off_bit_offsets, B + 1, ti64),
                               order=dv(lay.off_order, B, ti32), sum_offsets=dv(lay.off_sum_offsets, B + 1, ti64))
        if has_users:
            meta.user_summary = slot.users[:int(lay.us_elems)]
            meta.us_offsets, meta.cu_users = dv(lay.off_us_offsets, B + 1, ti64), dv(lay.off_cu_users, B + 1, ti32)
            meta.us_cols = dv(lay.off_us_cols, B, ti32)
        self.meta = meta
        self.features = slot.features[:T]
        self.cu_steps = dv(lay.off_cu_steps, B + 1, ti32)
        self.h2d_bytes = int(lay.blob_bytes) + T * slot.features.shape[1] * slot.features.element_size() + \
            int(lay.us_elems) * (slot.users.element_size() if has_users else 0)
        self.ready: Optional[torch.cuda.Event] = None
        self.collate_ms = self.issue_ms = 0.0

    def wait(self):
        """Make the current stream wait for this batch's host-to-device copies."""
        if self.ready is not None:
            torch.cuda.current_stream(self._slot.device).wait_event(self.ready)
        return self

    def release(self, event: Optional[torch.cuda.Event] = None):
        """The consumer is done with the batch once `event` (default: everything queued on the current stream so far)
        has completed; the loader may then overwrite the slot."""
        if self._slot is None:
            return
        if event is None:
            event = torch.cuda.Event()
            event.record(torch.cuda.current_stream(self._slot.device))
        slot, self._slot = self._slot, None
        slot.free_event = event
        slot.free.set()


class _Slot:
    def __init__(self, ds, device, blob_bytes, max_T, max_us, feat_dtype, us_dtype):
        self.ds, self.device = ds, device
        self.blob_host = torch.empty(blob_bytes, dtype=torch.uint8, pin_memory=True)
        self.blob_np = self.blob_host.numpy()
        self.blob_dev = torch.empty(blob_bytes, dtype=torch.uint8, device=device)
        self.features = torch.empty((max_T, ds.feature_dim), dtype=feat_dtype, device=device)
        self.users = torch.empty(max(max_us, 1), dtype=us_dtype, device=device)
        self.free = threading.Event()
        self.free.set()
        self.free_event: Optional[torch.cuda.Event] = None


class PackedEvalLoader:
    """Iterates `EvalBatch`es over a `PackedDataset(split="val", resident="pinned")` in dataset order (the reference's
    val loader does not shuffle).  `cycle=True` restarts from the first batch for ever (benchmarks)."""

    def __init__(self, dataset: PackedDataset, batch_size: int, device=None, slots: int = 3, cycle: bool = False,
                 with_users: bool = True):
        if dataset.resident != "pinned":
            raise _cabi.VsumError("PackedEvalLoader needs PackedDataset(..., resident='pinned'): batches are DMA'd out of the dataset")
        if not torch.cuda.is_available():
            raise _cabi.VsumError("PackedEvalLoader needs a CUDA device (no CPU fallback)")
        self.ds, self.bs, self.cycle, self.with_users = dataset, int(batch_size), cycle, with_users
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._L = _cabi.load()
        n = len(dataset)
        self._batches = [list(range(s, min(s + self.bs, n))) for s in range(0, n, self.bs)]
        lays = [self._layout(b) for b in self._batches]
        self.max_T = max((int(l.T) for l in lays), default=0)
        max_us = max((int(l.us_elems) for l in lays), default=0) if with_users else 0
        blob = max((int(l.blob_bytes) for l in lays), default=256)
        u8 = any(l.user_summary_dtype == _cabi.USER_SUMMARY_U8 for l in lays)
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.Stream(self.device)
            self._slots = [_Slot(dataset, self.device, blob, self.max_T, max_us,
                                 torch.bfloat16 if dataset.features_bf16 else torch.float32,
                                 torch.uint8 if u8 else torch.float32) for _ in range(max(2, int(slots)))]
        self.collate_ms: List[float] = []
        self.issue_ms: List[float] = []

    def __len__(self):
        return len(self._batches)

    def _raw(self, ids):
        return np.asarray([self.ds.ids[i] for i in ids], dtype=np.int32)

    def _layout(self, ids) -> _cabi.EvalBatchLayout:
        lay, raw = _cabi.EvalBatchLayout(), self._raw(ids)
        _cabi.check(self._L.vsum_pack_eval_collate(self.ds._h, raw.ctypes.data, len(ids), None, 0, C.byref(lay)), "vsum_pack_eval_collate")
        return lay

    def _prepare(self, ids, slot: _Slot) -> EvalBatch:
        t0 = time.perf_counter()
        lay, raw = _cabi.EvalBatchLayout(), self._raw(ids)
        _cabi.check(self._L.vsum_pack_eval_collate(self.ds._h, raw.ctypes.data, len(ids), slot.blob_host.data_ptr(),
                                                   slot.blob_host.numel(), C.byref(lay)), "vsum_pack_eval_collate")
        b = EvalBatch(slot, lay, ids, self.ds.names)
        b.collate_ms = (time.perf_counter() - t0) * 1e3          # host time of the metadata gather
        t1 = time.perf_counter()
        with torch.cuda.stream(self._stream):
            if slot.free_event is not None:
                self._stream.wait_event(slot.free_event)          # the previous batch of this slot has been consumed
                slot.free_event = None
            nb = int(lay.blob_bytes)
            slot.blob_dev[:nb].copy_(slot.blob_host[:nb], non_blocking=True)
            users = b.meta.user_summary.data_ptr() if (self.with_users and b.meta.has_users) else None
            _cabi.check(self._L.vsum_pack_h2d(self.ds._h, slot.blob_host.data_ptr(), C.byref(lay), b.features.data_ptr(), users,
                                              self._stream.cuda_stream), "vsum_pack_h2d")
            b.ready = torch.cuda.Event()
            b.ready.record(self._stream)
        b.issue_ms = (time.perf_counter() - t1) * 1e3            # issuing the copies (blocks while the stream's queue is full)
        return b

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=max(1, len(self._slots) - 1))
        stop = threading.Event()
        for slot in self._slots:          # slots of batches a previous (closed) iteration prepared but never handed out:
            if not slot.free.is_set():    # their copies sit on the loader's stream, ahead of anything issued from now on
                slot.free.set()

        def worker():
            try:
                torch.cuda.set_device(self.device)
                k = 0
                while True:
                    for ids in self._batches:
                        slot = self._slots[k % len(self._slots)]
                        while not slot.free.wait(timeout=0.05):
                            if stop.is_set():
                                return
                        if stop.is_set():
                            return
                        slot.free.clear()
                        q.put(self._prepare(ids, slot))
                        k += 1
                    if not self.cycle:
                        break
                q.put(None)
            except BaseException as e:                               # surface loader errors in the consumer
                q.put(e)

        t = threading.Thread(target=worker, daemon=True, name="vsum-eval-loader")
        t.start()
        prev: Optional[EvalBatch] = None
        try:
            while True:
                item = q.get()
                if prev is not None:
                    prev.release()                                   # not released explicitly: everything queued so far
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                self.collate_ms.append(item.collate_ms)
                self.issue_ms.append(item.issue_ms)
                prev = item
                yield item.wait()
        finally:
            stop.set()
            if prev is not None:
                prev.release()
            while t.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                t.join(timeout=0.05)
