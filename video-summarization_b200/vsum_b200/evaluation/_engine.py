"""Host side of the evaluation half of the hot path: packs per-video metadata (change points,
picks, user summaries) into flat arrays, moves them to the GPU and drives the three C-ABI calls

    vsum_shot_mean -> vsum_knapsack -> vsum_summary_fscore

that replace `generate_summary` + `knapSack` + `evaluate_summary` of the reference
(`src/evaluation/generate_summary.py:6-57`, `knapsack_implementation.py:1-30`,
`evaluation_metrics.py:4-33`).  No arithmetic of the path happens here: the host only
concatenates arrays and computes buffer sizes.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from .. import _cabi


def _device(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _cabi.VsumError("vsum_b200.evaluation needs a CUDA device: the B200 kernels have no CPU fallback")
    return torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())


def _cu(counts) -> np.ndarray:
    out = np.zeros(len(counts) + 1, dtype=np.int64)
    np.cumsum(np.asarray(counts, dtype=np.int64), out=out[1:])
    return out


# Knapsack kernel classes (vsum_knapsack_class_width): a video runs in the smallest class whose width
# holds capacity + 1 cells; one launch per class, so short videos never pin a whole SM's shared
# memory while the scorer of the next batch is running.  Mirrors kKnapsackClassWidth in vsum_eval.cu
# (tests/test_cabi_symbols.py checks the two tables agree).
KNAPSACK_CLASS_WIDTHS = np.array([256, 1024, 4096, 9728, 18944, 28672], dtype=np.int64)


def knapsack_class_width(caps: np.ndarray) -> np.ndarray:
    idx = np.searchsorted(KNAPSACK_CLASS_WIDTHS, np.asarray(caps, dtype=np.int64) + 1, side="left")
    if np.any(idx >= len(KNAPSACK_CLASS_WIDTHS)):
        raise _cabi.VsumError(f"knapsack capacity {int(np.max(caps))} exceeds the largest kernel class "
                              f"({int(KNAPSACK_CLASS_WIDTHS[-1])} cells: n_frames <= 191146)")
    return KNAPSACK_CLASS_WIDTHS[idx]


def capacity_of(last_end: int) -> int:
    """Buffer sizing only (the kernel computes the capacity itself, generate_summary.py:45-46)."""
    return int((int(last_end) + 1) * 0.15)


@dataclass
class HostEvalBatch:
    """Flat host arrays for B videos.  Everything here is input data or a buffer size."""
    B: int
    picks: np.ndarray            # int32[sum N_picks]
    cu_picks: np.ndarray         # int32[B+1]
    n_frames: np.ndarray         # int32[B]
    cps: np.ndarray              # int32[S_total,2]
    cu_shots: np.ndarray         # int32[B+1]
    bit_offsets: np.ndarray      # int64[B+1]  word offsets of each video's decision-bit matrix
    order: np.ndarray            # int32[B]    largest capacity first
    max_cap: int
    launches: list               # [(first index into order, count, max capacity)] one per kernel class
    sum_offsets: np.ndarray      # int64[B+1]  summary v occupies [sum_offsets[v], sum_offsets[v+1])
    user_summary: Optional[np.ndarray]   # flat; float32, or uint8 when every video's rows came as uint8 / bool
    us_offsets: Optional[np.ndarray]     # int64[B+1]
    cu_users: Optional[np.ndarray]       # int32[B+1]
    us_cols: Optional[np.ndarray]        # int32[B]

    @staticmethod
    def build(change_points: Sequence[np.ndarray], n_frames: Sequence, picks: Sequence[np.ndarray],
              user_summaries: Optional[Sequence[np.ndarray]] = None) -> "HostEvalBatch":
        B = len(change_points)
        cps_list = [np.ascontiguousarray(np.asarray(c).reshape(-1, 2), dtype=np.int32) for c in change_points]
        pk_list = []
        for p in picks:
            p = np.asarray(p)
            pk_list.append(p.astype(np.int32) if p.dtype != np.int32 else p)     # generate_summary.py:27-28
        n_shots = [len(c) for c in cps_list]
        last_end = [int(c[-1, 1]) if len(c) else -1 for c in cps_list]
        caps = [capacity_of(e) if e >= 0 else 0 for e in last_end]
        caps_np = np.asarray(caps, dtype=np.int64)
        widths = knapsack_class_width(caps_np) if B else np.zeros(0, np.int64)
        words = [int(s) * int(w // 32) for s, w in zip(n_shots, widths)]
        order = np.argsort(-caps_np, kind="stable").astype(np.int32)
        launches, pos = [], 0
        while pos < B:
            w = widths[order[pos]]
            end = pos
            while end < B and widths[order[end]] == w:
                end += 1
            launches.append((pos, end - pos, int(caps_np[order[pos]])))
            pos = end
        us = us_off = cu_users = us_cols = None
        if user_summaries is not None:
            mats, us_dt = _user_matrices(user_summaries)
            us = np.concatenate([m.reshape(-1) for m in mats]) if B else np.zeros(0, us_dt)
            us_off = _cu([m.size for m in mats])
            cu_users = _cu([m.shape[0] for m in mats]).astype(np.int32)
            us_cols = np.asarray([m.shape[1] for m in mats], dtype=np.int32)
        return HostEvalBatch(
            B=B,
            picks=np.concatenate(pk_list).astype(np.int32, copy=False) if B else np.zeros(0, np.int32),
            cu_picks=_cu([len(p) for p in pk_list]).astype(np.int32),
            n_frames=np.asarray([int(np.asarray(n)) for n in n_frames], dtype=np.int32),
            cps=np.concatenate(cps_list) if B else np.zeros((0, 2), np.int32),
            cu_shots=_cu(n_shots).astype(np.int32),
            bit_offsets=_cu(words),
            order=order,
            max_cap=max(caps) if B else 0,
            launches=launches,
            sum_offsets=_cu([e + 1 for e in last_end]),
            user_summary=us, us_offsets=us_off, cu_users=cu_users, us_cols=us_cols)


def _user_matrices(user_summaries):
    """[U, n_frames] matrices of one dtype for the overlap kernel: uint8 when every video's user summary is stored as
    uint8 / bool (the packed dataset's lossless form of the 0/1 rows), else float32 as the h5 files hold them."""
    arrs = [np.asarray(u) for u in user_summaries]
    dt = np.uint8 if arrs and all(a.dtype in (np.uint8, np.bool_) for a in arrs) else np.float32
    return [np.ascontiguousarray(a, dtype=dt).reshape(len(a), -1) for a in arrs], dt


def _us_dtype_code(a) -> int:
    return _cabi.USER_SUMMARY_U8 if a is not None and a.dtype == np.uint8 else _cabi.USER_SUMMARY_F32


_FIELDS = ("picks", "cu_picks", "n_frames", "cps", "cu_shots", "bit_offsets", "order", "sum_offsets")
_USER_FIELDS = ("user_summary", "us_offsets", "cu_users", "us_cols")


class DeviceEvalBatch:
    """The same arrays resident in HBM.  `refill` re-uploads a (same-shaped) host batch into the
    existing device buffers from pinned staging copies, for the pipelined end-to-end path."""

    def __init__(self, hb: HostEvalBatch, device=None, pin: bool = False):
        self.host = hb
        self.device = dev = _device(device)
        self.has_users = hb.user_summary is not None
        self._names = _FIELDS + (_USER_FIELDS if self.has_users else ())
        self._staging = {}
        for name in self._names:
            t = torch.from_numpy(np.ascontiguousarray(getattr(hb, name)))
            if pin:
                t = t.pin_memory()
                self._staging[name] = t
            setattr(self, name, t.to(dev, non_blocking=True))
        self.h2d_bytes = sum(getattr(hb, name).nbytes for name in self._names)

    def refill(self) -> None:
        """Host -> device copy of every array again (pinned staging, current stream)."""
        if not self._staging:
            raise ValueError("refill() needs a batch built with pin=True")
        for name in self._names:
            getattr(self, name).copy_(self._staging[name], non_blocking=True)


_scratch: dict = {}


def _scratch_buf(key: str, nbytes: int, dev: torch.device) -> torch.Tensor:
    """Per (purpose, device, STREAM) scratch: two streams (the pipelined side stream and a caller's main stream, or two
    Summarizers) never share knapsack decision bits / overlap counts, and a buffer that is outgrown is recorded on its
    stream before it is dropped."""
    stream = torch.cuda.current_stream(dev)
    k = (key, dev.index, stream.cuda_stream)
    t = _scratch.get(k)
    if t is None or t.numel() < nbytes:
        if t is not None:
            t.record_stream(stream)
        t = torch.empty(max(int(nbytes * 1.25), 1024), dtype=torch.uint8, device=dev)
        _scratch[k] = t
    return t


_aux: dict = {}


def _aux_streams(dev: torch.device, cur, n: int):
    """`n` helper streams per (device, calling stream) for the knapsack class launches."""
    k = (dev.index, cur.cuda_stream)
    pool = _aux.setdefault(k, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(dev))
    return pool[:n]


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def summarize(db: DeviceEvalBatch, scores: torch.Tensor, cu_steps: torch.Tensor, method: str = "avg",
              want_f: bool = True, want_per_user: bool = False) -> dict:
    """scores float32[T] on db.device, cu_steps int32[B+1] on db.device.  Returns device tensors:
    val fp64[S], wt int32[S], cap int32[B], selected uint8[S], summary int8[sum], f fp64[B]."""
    hb, dev = db.host, db.device
    L = _cabi.load()
    if method not in ("avg", "max"):
        method = "avg"          # evaluation_metrics.py:30-33: anything but 'max' averages
    B, S = hb.B, int(hb.cu_shots[-1])
    if scores.dtype != torch.float32 or not scores.is_contiguous() or scores.device != dev:
        raise ValueError("scores must be a contiguous float32 tensor on the batch's device")
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        val = torch.empty(S, dtype=torch.float64, device=dev)
        wt = torch.empty(S, dtype=torch.int32, device=dev)
        cap = torch.empty(B, dtype=torch.int32, device=dev)
        selected = torch.empty(S, dtype=torch.uint8, device=dev)
        summary = torch.empty(int(hb.sum_offsets[-1]), dtype=torch.int8, device=dev)
        out = dict(val=val, wt=wt, cap=cap, selected=selected, summary=summary, f=None, per_user=None)
        if B == 0:
            return out
        _cabi.check(L.vsum_shot_mean(scores.data_ptr(), cu_steps.data_ptr(), db.picks.data_ptr(),
                                     db.cu_picks.data_ptr(), db.n_frames.data_ptr(), db.cps.data_ptr(),
                                     db.cu_shots.data_ptr(), B, S, val.data_ptr(), wt.data_ptr(),
                                     cap.data_ptr(), stream), "vsum_shot_mean")
        bits = _scratch_buf("bits", int(hb.bit_offsets[-1]) * 4, dev)
        # One launch per capacity class; the classes are independent (disjoint videos), so they run side by side on forked
        # streams instead of one after the other: the widest class keeps only as many SMs busy as it has videos.
        cur = torch.cuda.current_stream(dev)
        launches = [l for l in hb.launches if l[1] > 0]
        aux = _aux_streams(dev, cur, len(launches) - 1)
        fork = torch.cuda.Event()
        fork.record(cur)
        for i, (first, count, max_cap) in enumerate(launches):
            st = cur if i == 0 else aux[i - 1]
            if i:
                st.wait_event(fork)
            _cabi.check(L.vsum_knapsack(val.data_ptr(), wt.data_ptr(), db.cu_shots.data_ptr(), cap.data_ptr(),
                                        db.bit_offsets.data_ptr(), db.order.data_ptr() + 4 * first, count, max_cap,
                                        bits.data_ptr(), selected.data_ptr(), st.cuda_stream), "vsum_knapsack")
            if i:
                cur.wait_stream(st)
        f = per_user = counts = None
        total_users = 0
        if want_f:
            if not db.has_users:
                raise ValueError("F-score requested but the batch was built without user summaries")
            total_users = int(hb.cu_users[-1])
            f = torch.empty(B, dtype=torch.float64, device=dev)
            counts = _scratch_buf("counts", max(total_users, 1) * 24, dev)
            if want_per_user:
                per_user = torch.empty(total_users, dtype=torch.float64, device=dev)
        _cabi.check(L.vsum_summary_fscore(
            selected.data_ptr(), db.cps.data_ptr(), db.cu_shots.data_ptr(),
            _ptr(db.user_summary) if want_f else None, _us_dtype_code(hb.user_summary),
            _ptr(db.us_offsets) if want_f else None,
            _ptr(db.cu_users) if want_f else None, _ptr(db.us_cols) if want_f else None,
            B, total_users, _cabi.FSCORE_MAX if method == "max" else _cabi.FSCORE_AVG,
            summary.data_ptr(), db.sum_offsets.data_ptr(), int(hb.sum_offsets[-1]), _ptr(counts),
            _ptr(f), _ptr(per_user), stream), "vsum_summary_fscore")
        out["f"], out["per_user"] = f, per_user
    return out


def fscore_of_masks(summaries: Sequence[np.ndarray], user_summaries: Sequence[np.ndarray], method: str,
                    device=None) -> np.ndarray:
    """`evaluate_summary` for arbitrary 0/1 masks (selected == NULL path of vsum_summary_fscore)."""
    dev = _device(device)
    L = _cabi.load()
    B = len(summaries)
    masks = [np.ascontiguousarray(np.asarray(s), dtype=np.int8) for s in summaries]
    mats, us_dt = _user_matrices(user_summaries)
    sum_off = _cu([len(m) for m in masks])
    us_off = _cu([m.size for m in mats])
    cu_users = _cu([m.shape[0] for m in mats]).astype(np.int32)
    us_cols = np.asarray([m.shape[1] for m in mats], dtype=np.int32)
    total_users = int(cu_users[-1])
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_mask = t(np.concatenate(masks) if B else np.zeros(0, np.int8))
        d_us = t(np.concatenate([m.reshape(-1) for m in mats]) if B else np.zeros(0, us_dt))
        d_sum_off, d_us_off, d_cu_users, d_cols = t(sum_off), t(us_off), t(cu_users), t(us_cols)
        f = torch.empty(B, dtype=torch.float64, device=dev)
        counts = _scratch_buf("counts", max(total_users, 1) * 24, dev)
        _cabi.check(L.vsum_summary_fscore(
            None, None, None, d_us.data_ptr(), _cabi.USER_SUMMARY_U8 if us_dt is np.uint8 else _cabi.USER_SUMMARY_F32,
            d_us_off.data_ptr(), d_cu_users.data_ptr(), d_cols.data_ptr(),
            B, total_users, _cabi.FSCORE_MAX if method == "max" else _cabi.FSCORE_AVG,
            d_mask.data_ptr(), d_sum_off.data_ptr(), int(sum_off[-1]), counts.data_ptr(), f.data_ptr(), None,
            stream), "vsum_summary_fscore")
        return f.cpu().numpy()


def rank_correlations(scores: Sequence[np.ndarray], picks: Sequence[np.ndarray], n_frames: Sequence,
                      user_scores: Sequence[np.ndarray], device=None, per_user: bool = False):
    """Kendall tau-b / Spearman rho of every video's predicted frame scores against each user's frame
    scores (compute_correlation.py:4-15) in one batched GPU pass.  `scores[v]` are the sub-sampled
    scores (float32[N_v]); the upsampling of compute_metrics.py:19-39 happens inside the kernels.
    Returns (kendall[B], spearman[B]) -- the per-video `sum(x) / len(x)` over users (line 15) -- and,
    with `per_user`, also the two flat fp64 per-user arrays.

    Result dtype follows scipy's: with float32 user scores the rank arrays are float32, so
    `stats.kendalltau` hands back its fp64 statistic rounded to float32 and line 15 sums float32
    scalars; any other dtype keeps fp64.  `stats.spearmanr` is fp64 either way.  The kernels always
    produce the fp64 statistic (tau bit-exact with scipy's internal value); the rounding and the
    per-video mean over a handful of users are applied here."""
    dev = _device(device)
    B = len(scores)
    if B == 0:
        z = np.zeros(0, np.float64)
        return (z, z, z, z) if per_user else (z, z)
    sc = [np.ascontiguousarray(np.asarray(s), dtype=np.float32).reshape(-1) for s in scores]
    pk = [np.asarray(p).astype(np.int32, copy=False).reshape(-1) for p in picks]
    for s_, p_ in zip(sc, pk):
        if len(s_) != len(p_):
            raise ValueError("every video needs one pick per score")
    raw = [np.asarray(u) for u in user_scores]
    tau_dtypes = [np.float32 if r.dtype in (np.float32, np.float16) else np.float64 for r in raw]
    mats = [np.ascontiguousarray(r, dtype=np.float32).reshape(len(r), -1) for r in raw]
    cu_steps = _cu([len(s_) for s_ in sc]).astype(np.int32)
    us_off = _cu([m.size for m in mats])
    cu_users = _cu([m.shape[0] for m in mats]).astype(np.int32)
    us_cols = np.asarray([m.shape[1] for m in mats], dtype=np.int32)
    nfr = np.asarray([int(np.asarray(n)) for n in n_frames], dtype=np.int32)
    T, total_users, total_elems = int(cu_steps[-1]), int(cu_users[-1]), int(us_off[-1])
    max_steps = max(len(s_) for s_ in sc)
    with torch.cuda.device(dev):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        d_sc = t(np.concatenate(sc)) if T else torch.zeros(1, device=dev)
        d_pk = t(np.concatenate(pk)) if T else torch.zeros(1, dtype=torch.int32, device=dev)
        d_us = t(np.concatenate([m.reshape(-1) for m in mats])) if total_elems else torch.zeros(1, device=dev)
        d_cu, d_nf, d_off, d_cuu, d_cols = t(cu_steps), t(nfr), t(us_off), t(cu_users), t(us_cols)
        L = _cabi.load()
        need = L.vsum_rank_correlation_workspace_bytes(total_elems, T, B, total_users)
        ws = _scratch_buf("corr", need + 1024, dev)
        wp = (ws.data_ptr() + 1023) // 1024 * 1024
        out = torch.empty(2 * B + 2 * max(total_users, 1), dtype=torch.float64, device=dev)
        tau, rho, pu_t, pu_r = out[:B], out[B:2 * B], out[2 * B:2 * B + max(total_users, 1)], out[2 * B + max(total_users, 1):]
        _cabi.check(L.vsum_rank_correlation(
            d_sc.data_ptr(), d_cu.data_ptr(), d_pk.data_ptr(), d_nf.data_ptr(), d_us.data_ptr(), d_off.data_ptr(),
            d_cuu.data_ptr(), d_cols.data_ptr(), B, T, max_steps, total_users, total_elems, wp,
            ws.numel() - (wp - ws.data_ptr()), tau.data_ptr(), rho.data_ptr(), pu_t.data_ptr(), pu_r.data_ptr(),
            torch.cuda.current_stream(dev).cuda_stream), "vsum_rank_correlation")
        host = out.cpu().numpy()
    pu_tau, pu_rho = host[2 * B:2 * B + total_users], host[2 * B + max(total_users, 1):2 * B + max(total_users, 1) + total_users]
    taus, rhos = [], []
    for v in range(B):                                   # line 15: sum(kendal) / len(kendal), in scipy's result dtype
        u0, u1 = int(cu_users[v]), int(cu_users[v + 1])
        if u1 == u0:
            raise ZeroDivisionError("division by zero")  # what line 15 does for a video without users
        taus.append(sum(list(pu_tau[u0:u1].astype(tau_dtypes[v]))) / (u1 - u0))
        rhos.append(sum(list(pu_rho[u0:u1])) / (u1 - u0))
    res = (taus, rhos)
    if per_user:
        res += (pu_tau.copy(), pu_rho.copy())
    return res
