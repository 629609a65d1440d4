"""Packed dataset file + loader.

* `write_pack(path, videos)` / `convert_h5(h5_path, out)` produce one `.vspack` file per dataset: a 64-byte
  header, 4096-byte aligned raw arrays and a fixed-size index (layout in `csrc/vsum_io.cu`).
* `PackedDataset` maps the file through the native reader (`vsum_pack_*`, zero-copy numpy views) and yields
  what `TSDataset.__getitem__` yields (`src/data/dataset.py:127-135`): `(features, targets)` for the train
  split, `(features, targets, UserSummaries)` for the val split.
* `PackedLoader` replaces DataLoader + `collate_fn_train` / `collate_fn_pretrain` (`dataset.py:139-168`,
  `train.py:58-72`): a background thread runs the native multi-threaded collate into double-buffered PINNED
  staging, the copy to the GPU is issued on its own stream, and every batch arrives as packed rows
  `[sum N, 1024]` + `cu_seqlens` -- no `pad_sequence`, no 1000-sentinel mask (`train.py:115-118`).
"""
from __future__ import annotations

import ctypes as C
import os
import queue
import struct
import threading
from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np
import torch

from .. import _cabi

MAGIC = b"VSPACK01"
_ALIGN = 4096
_INDEX_ENTRY = 256


class UserSummaries:
    """Same record as `src/data/dataset.py:146-154`."""

    def __init__(self, user_summary, user_scores, name, changes_point, n_frames, picks):
        self.user_summary = user_summary
        self.user_scores = user_scores
        self.change_points = changes_point
        self.n_frames = n_frames
        self.picks = picks
        self.name = name


def _pad_to(f, align):
    pos = f.tell()
    pad = (-pos) % align
    if pad:
        f.write(b"\0" * pad)
    return pos + pad


def to_bf16_bits(a: np.ndarray) -> np.ndarray:
    """float32 -> bfloat16 bit patterns (uint16), round to nearest even -- the rounding `__float2bfloat16_rn` and
    torch's `.bfloat16()` apply (NaN payloads aside; features are finite)."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    return ((u + (np.uint32(0x7fff) + ((u >> np.uint32(16)) & np.uint32(1)))) >> np.uint32(16)).astype(np.uint16)


def write_pack(path: str, videos: Iterable[dict], feature_dim: int = 1024, user_summary_u8: bool = False,
               features_bf16: bool = False) -> None:
    """`videos`: dicts with `name`, `features` f32[N,dim], and optionally `gtscore` f32[N], `picks` int[N],
    `change_points` int[S,2], `n_frames`, `user_summary` [U,n_frames], `user_scores` [U,n_frames],
    `video_rep` f32[rep_dim] (the pretraining target, dataset.py:26).  `user_summary_u8` stores the 0/1
    user summaries as bytes (4x smaller than the h5 files' float32).  `features_bf16` stores the features rounded to
    bfloat16 (half the bytes on disk, over PCIe and out of HBM; `features` may also arrive already rounded, as uint16
    bfloat16 bit patterns): an inference-side option -- the bf16 scorer then runs
    its feature GEMM in bf16 (`VSUM_MODE_BF16_FEATURES`); training packs keep float32."""
    entries = []
    with open(path, "wb") as f:
        f.write(b"\0" * 64)
        for v in videos:
            pre_rounded = features_bf16 and getattr(v["features"], "dtype", None) == np.uint16     # bfloat16 bit patterns as given
            feats = np.ascontiguousarray(v["features"]) if pre_rounded else np.ascontiguousarray(v["features"], dtype=np.float32)
            if feats.ndim != 2 or feats.shape[1] != feature_dim:
                raise ValueError(f"{v.get('name')}: features must be [N,{feature_dim}]")
            n = feats.shape[0]
            off = [0] * 7
            def put(kind, arr, align=64):
                off[kind] = _pad_to(f, align)
                f.write(memoryview(np.ascontiguousarray(arr)).cast("B"))
            put(_cabi.PACK_FEATURES, feats if (pre_rounded or not features_bf16) else to_bf16_bits(feats), _ALIGN)
            n_frames = n_shots = n_users = rep_dim = has_scores = 0
            if v.get("gtscore") is not None:
                g = np.ascontiguousarray(v["gtscore"], dtype=np.float32).reshape(-1)
                if len(g) != n:
                    raise ValueError("gtscore needs one value per step")
                put(_cabi.PACK_GTSCORE, g)
            if v.get("picks") is not None:
                pk = np.ascontiguousarray(v["picks"]).astype(np.int32).reshape(-1)
                if len(pk) != n:
                    raise ValueError("picks needs one value per step")
                put(_cabi.PACK_PICKS, pk)
            if v.get("change_points") is not None:
                cps = np.ascontiguousarray(v["change_points"]).astype(np.int32).reshape(-1, 2)
                n_shots = len(cps)
                put(_cabi.PACK_CHANGE_POINTS, cps)
            if v.get("n_frames") is not None:
                n_frames = int(np.asarray(v["n_frames"]))
            if v.get("user_summary") is not None:
                us = np.asarray(v["user_summary"])
                us = us.reshape(len(us), -1)
                n_users, n_frames = us.shape[0], n_frames or us.shape[1]
                if us.shape[1] != n_frames:
                    raise ValueError("user_summary rows must have n_frames columns")
                if user_summary_u8 and not np.array_equal(us, us.astype(np.uint8)):
                    raise ValueError("user_summary_u8 needs integer values in [0, 255] (the byte form must be lossless)")
                put(_cabi.PACK_USER_SUMMARY, us.astype(np.uint8 if user_summary_u8 else np.float32), _ALIGN)
            if v.get("user_scores") is not None:
                sc = np.asarray(v["user_scores"], dtype=np.float32)
                sc = sc.reshape(len(sc), -1)
                if sc.shape != (n_users, n_frames):
                    raise ValueError("user_scores must match user_summary's shape")
                has_scores = 1
                put(_cabi.PACK_USER_SCORES, sc, _ALIGN)
            if v.get("video_rep") is not None:
                rep = np.ascontiguousarray(v["video_rep"], dtype=np.float32).reshape(-1)
                rep_dim = len(rep)
                put(_cabi.PACK_VIDEO_REP, rep)
            name = str(v.get("name", f"video_{len(entries)}")).encode()[:95]
            entries.append(struct.pack("<96s8i7Q", name, n, n_frames, n_shots, n_users, rep_dim, has_scores,
                                       1 if user_summary_u8 else 0, 0, *off).ljust(_INDEX_ENTRY, b"\0"))
        index_offset = _pad_to(f, _ALIGN)
        for e in entries:
            f.write(e)
        total = f.tell()
        f.seek(0)
        f.write(struct.pack("<8sIIQQII", MAGIC, 1, len(entries), index_offset, total, feature_dim,
                            _cabi.FEATURES_BF16 if features_bf16 else _cabi.FEATURES_F32).ljust(64, b"\0"))


def convert_h5(h5_path: str, out_path: str, video_rep_dir: Optional[str] = None, **kw) -> None:
    """One-off conversion of a DSNet-style h5 file (`dataset.py:89-103`) into a pack file.  Needs h5py, which
    is only required here -- training and evaluation read the pack."""
    try:
        import h5py
    except ImportError as e:                                   # not in this image; conversion runs where the data lives
        raise ImportError("convert_h5 needs h5py; run the conversion on a machine that has it") from e

    def gen():
        with h5py.File(h5_path, "r") as f:
            for key in f.keys():
                g = f[key]
                rec = dict(name=key, features=g["features"][...].astype(np.float32))
                for src, dst in (("gtscore", "gtscore"), ("picks", "picks"), ("change_points", "change_points"),
                                 ("n_frames", "n_frames"), ("user_summary", "user_summary"), ("user_scores", "user_scores")):
                    if src in g:
                        rec[dst] = np.array(g[src])
                if video_rep_dir:
                    rec["video_rep"] = np.load(os.path.join(video_rep_dir, f"{key}.npy"))
                yield rec
    write_pack(out_path, gen(), **kw)


_NP = {_cabi.PACK_FEATURES: np.float32, _cabi.PACK_GTSCORE: np.float32, _cabi.PACK_PICKS: np.int32,
       _cabi.PACK_CHANGE_POINTS: np.int32, _cabi.PACK_USER_SCORES: np.float32, _cabi.PACK_VIDEO_REP: np.float32}


class PackedDataset(torch.utils.data.Dataset):
    """`split="train"` -> `(features, targets)`; `split="val"` -> `(features, targets, UserSummaries)`;
    `split="pretrain"` -> `(features, video_rep)`  (dataset.py:33-37, 127-135)."""

    def __init__(self, path: str, split: str = "train", keys: Optional[Sequence[str]] = None, min_steps: int = 0,
                 resident: str = "mmap"):
        """`resident="pinned"`: the file is read once into page-locked host memory, so that `PackedEvalLoader` can DMA
        every batch straight out of the dataset (no host-side gather of the features / user summaries)."""
        self.path, self.split = path, split
        if resident not in ("mmap", "pinned"):
            raise ValueError("resident must be 'mmap' or 'pinned'")
        L = _cabi.load()
        h = C.c_void_p()
        _cabi.check(L.vsum_pack_open_ex(os.fsencode(path), _cabi.PACK_PINNED if resident == "pinned" else _cabi.PACK_MMAP,
                                        C.byref(h)), "vsum_pack_open_ex")
        self.resident = resident
        self._h, self._L = h, L
        self.feature_dim = int(L.vsum_pack_feature_dim(h))
        self.features_bf16 = int(L.vsum_pack_feature_dtype(h)) == _cabi.FEATURES_BF16
        if self.features_bf16 and split != "val":
            raise ValueError("a pack written with features_bf16 serves inference (split='val'); training reads float32 features")
        self.info: List[_cabi.PackInfo] = []
        for i in range(int(L.vsum_pack_num_videos(h))):
            inf = _cabi.PackInfo()
            _cabi.check(L.vsum_pack_video_info(h, i, C.byref(inf)), "vsum_pack_video_info")
            self.info.append(inf)
        names = [inf.name.decode() for inf in self.info]
        want = None if keys is None else {os.path.basename(str(k)) for k in keys}        # dataset.py:137-140
        self.ids = [i for i, nme in enumerate(names)
                    if (want is None or nme in want) and self.info[i].n_steps > min_steps]   # dataset.py:121 uses > 50
        self.names = [names[i] for i in self.ids]

    def close(self):
        if getattr(self, "_h", None):
            self._L.vsum_pack_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return len(self.ids)

    def array(self, idx: int, kind: int) -> Optional[np.ndarray]:
        """Zero-copy (read-only) numpy view of one array of the idx-th selected video."""
        i = self.ids[idx]
        ptr, nbytes = C.c_void_p(), C.c_uint64()
        _cabi.check(self._L.vsum_pack_array(self._h, i, kind, C.byref(ptr), C.byref(nbytes)), "vsum_pack_array")
        if not ptr.value:
            return None
        inf = self.info[i]
        dt = _NP.get(kind) or (np.uint8 if inf.user_summary_dtype == 1 else np.float32)
        if kind == _cabi.PACK_FEATURES and self.features_bf16:
            dt = np.uint16                                      # bfloat16 bit patterns (numpy has no bf16)
        buf = (C.c_uint8 * nbytes.value).from_address(ptr.value)
        a = np.frombuffer(buf, dtype=dt)
        a.flags.writeable = False
        shape = {_cabi.PACK_FEATURES: (inf.n_steps, self.feature_dim), _cabi.PACK_CHANGE_POINTS: (inf.n_shots, 2),
                 _cabi.PACK_USER_SUMMARY: (inf.n_users, inf.n_frames), _cabi.PACK_USER_SCORES: (inf.n_users, inf.n_frames)}.get(kind)
        return a.reshape(shape) if shape else a

    def user(self, idx: int) -> UserSummaries:
        inf = self.info[self.ids[idx]]
        return UserSummaries(self.array(idx, _cabi.PACK_USER_SUMMARY), self.array(idx, _cabi.PACK_USER_SCORES), self.names[idx],
                             self.array(idx, _cabi.PACK_CHANGE_POINTS), np.array(inf.n_frames), self.array(idx, _cabi.PACK_PICKS))

    def n_steps(self, idx: int) -> int:
        return int(self.info[self.ids[idx]].n_steps)

    def __getitem__(self, idx):
        feats = torch.from_numpy(np.array(self.array(idx, _cabi.PACK_FEATURES)))          # owned copy, like dataset.py:131-134
        if self.features_bf16:
            feats = feats.view(torch.bfloat16)
        if self.split == "pretrain":
            return feats, torch.from_numpy(np.array(self.array(idx, _cabi.PACK_VIDEO_REP)))
        targets = torch.from_numpy(np.array(self.array(idx, _cabi.PACK_GTSCORE)))
        if self.split == "train":
            return feats, targets
        return feats, targets, self.user(idx)


@dataclass
class PackedBatch:
    """One collated batch: packed rows on `device` plus the host-side lengths."""
    ids: List[int]                     # dataset indices, batch order
    seqlens: List[int]
    features: torch.Tensor             # [sum N, dim] on the device: fp32, or bf16 from a features_bf16 pack
    targets: Optional[torch.Tensor]    # [sum N] fp32 on the device (train / val)
    cu_seqlens: torch.Tensor           # int32[B+1] on the device
    video_rep: Optional[torch.Tensor]  # [B, rep_dim] (pretrain)
    ready: Optional[torch.cuda.Event] = None

    def wait(self):
        """Make the current stream wait for the batch's host-to-device copies.  The device tensors were allocated on the
        loader's copy stream: they are recorded on the consumer's stream, so the caching allocator cannot hand their blocks
        to the next copy while kernels queued by the consumer still read them."""
        if self.ready is not None:
            cur = torch.cuda.current_stream(self.features.device)
            cur.wait_event(self.ready)
            for t in (self.features, self.targets, self.cu_seqlens, self.video_rep):
                if t is not None and t.is_cuda:
                    t.record_stream(cur)
        return self


class PackedLoader:
    """Iterates `PackedBatch`es.  A worker thread gathers batch k+1 with the native multi-threaded collate into
    pinned staging and enqueues its copy on a side stream while the caller computes on batch k."""

    def __init__(self, dataset: PackedDataset, batch_size: int, shuffle: bool = False, device=None, seed: int = 0,
                 collate_threads: int = 8, drop_last: bool = False, prefetch: int = 2):
        self.ds, self.bs, self.shuffle, self.drop_last = dataset, int(batch_size), shuffle, drop_last
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))
        self.threads, self.prefetch = int(collate_threads), max(1, int(prefetch))
        self._rng = np.random.default_rng(seed)
        self._stream = torch.cuda.Stream(self.device) if self.device.type == "cuda" else None

    def __len__(self):
        n = len(self.ds)
        return n // self.bs if self.drop_last else (n + self.bs - 1) // self.bs

    def _batches(self):
        order = self._rng.permutation(len(self.ds)) if self.shuffle else np.arange(len(self.ds))
        for s in range(0, len(order), self.bs):
            ids = [int(i) for i in order[s:s + self.bs]]
            if len(ids) == self.bs or not self.drop_last:
                yield ids

    def _collate(self, ids: List[int]) -> PackedBatch:
        ds, L = self.ds, self.ds._L
        lens = [ds.n_steps(i) for i in ids]
        T, pin = sum(lens), self.device.type == "cuda"
        feats = torch.empty((T, ds.feature_dim), dtype=torch.bfloat16 if ds.features_bf16 else torch.float32, pin_memory=pin)
        with_t = ds.split != "pretrain"
        tgt = torch.empty(T, dtype=torch.float32, pin_memory=pin) if with_t else None
        cu = torch.empty(len(ids) + 1, dtype=torch.int32, pin_memory=pin)
        raw = np.asarray([ds.ids[i] for i in ids], dtype=np.int32)
        _cabi.check(L.vsum_pack_collate(ds._h, raw.ctypes.data, len(ids), self.threads, feats.data_ptr(),
                                        tgt.data_ptr() if with_t else None, cu.data_ptr()), "vsum_pack_collate")
        rep = None
        if ds.split == "pretrain":
            rep = torch.from_numpy(np.stack([ds.array(i, _cabi.PACK_VIDEO_REP) for i in ids]))
            rep = rep.pin_memory() if pin else rep
        ready = None
        if self._stream is not None:
            with torch.cuda.stream(self._stream):
                feats_d, cu_d = feats.to(self.device, non_blocking=True), cu.to(self.device, non_blocking=True)
                tgt_d = tgt.to(self.device, non_blocking=True) if with_t else None
                rep_d = rep.to(self.device, non_blocking=True) if rep is not None else None
                ready = torch.cuda.Event()
                ready.record(self._stream)
            keep = (feats, tgt, cu, rep)                       # pinned staging stays alive until the batch is dropped
        else:
            feats_d, cu_d, tgt_d, rep_d, keep = feats, cu, tgt, rep, None
        b = PackedBatch(ids, lens, feats_d, tgt_d, cu_d, rep_d, ready)
        b._staging = keep
        return b

    def __iter__(self):
        q: "queue.Queue" = queue.Queue(maxsize=self.prefetch)
        stop = threading.Event()

        def worker():
            try:
                if self.device.type == "cuda":
                    torch.cuda.set_device(self.device)
                for ids in self._batches():
                    if stop.is_set():
                        return
                    q.put(self._collate(ids))
                q.put(None)
            except BaseException as e:                          # surface loader errors in the consumer
                q.put(e)

        t = threading.Thread(target=worker, daemon=True)
        t.start()
        try:
            while True:
                item = q.get()
                if item is None:
                    return
                if isinstance(item, BaseException):
                    raise item
                yield item.wait()
        finally:
            stop.set()
            while t.is_alive():
                try:
                    q.get_nowait()
                except queue.Empty:
                    pass
                t.join(timeout=0.05)
