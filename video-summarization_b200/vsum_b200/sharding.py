"""Multi-GPU partitioning of the evaluation pass.

Videos are independent through scorer, pooling, knapsack and F-score, so they are sharded across
ranks with NO data-path collective; the single exchange is an all-gather of the per-video fp64
F-scores at the end.  The mean is taken on the host in the original video order
(`compute_metrics.py:92`), so it is bit-identical for every world size.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch
import torch.distributed as dist


def scorer_cost(n_steps: int, d: int = 256, layers: int = 4, in_features: int = 1024) -> float:
    """Algorithmic forward FLOPs of one video."""
    n = float(n_steps)
    return n * (2 * in_features * d + layers * (24 * d * d + 4 * n * d) + 2 * d)


def partition(n_steps: Sequence[int], world_size: int) -> List[List[int]]:
    """Greedy longest-first assignment of video indices to ranks, balanced by scorer FLOPs.
    Deterministic and independent of the calling rank."""
    order = sorted(range(len(n_steps)), key=lambda i: (-n_steps[i], i))
    load = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += scorer_cost(n_steps[i])
    return [sorted(s) for s in shards]


def gather_fscores(local_idx: Sequence[int], local_f: np.ndarray, n_total: int, group=None) -> np.ndarray:
    """All ranks contribute (index, F) pairs; every rank gets the full fp64 vector in video
    order.  Works with the nccl (GPU tensors) and gloo (CPU tensors) backends."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        out = np.full(n_total, np.nan, dtype=np.float64)
        out[np.asarray(local_idx, dtype=np.int64)] = local_f
        return out
    world = dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    count = torch.tensor([len(local_idx)], dtype=torch.int64, device=dev)
    counts = [torch.zeros_like(count) for _ in range(world)]
    dist.all_gather(counts, count, group=group)
    cap = int(max(int(c.item()) for c in counts))
    idx = torch.full((cap,), -1, dtype=torch.int64, device=dev)
    val = torch.zeros((cap,), dtype=torch.float64, device=dev)
    idx[:len(local_idx)] = torch.as_tensor(np.asarray(local_idx, dtype=np.int64), device=dev)
    val[:len(local_idx)] = torch.as_tensor(np.asarray(local_f, dtype=np.float64), device=dev)
    all_idx = [torch.empty_like(idx) for _ in range(world)]
    all_val = [torch.empty_like(val) for _ in range(world)]
    dist.all_gather(all_idx, idx, group=group)
    dist.all_gather(all_val, val, group=group)
    out = np.full(n_total, np.nan, dtype=np.float64)
    for i_t, v_t in zip(all_idx, all_val):
        i_np, v_np = i_t.cpu().numpy(), v_t.cpu().numpy()
        keep = i_np >= 0
        out[i_np[keep]] = v_np[keep]
    return out


# ---------------------------------------------------------------------------------------------
# data-parallel training (BASELINE config 3): one process per GPU, NCCL gradient all-reduce
# ---------------------------------------------------------------------------------------------
def global_loss_denominator(local_batch: int, local_nmax: int, group=None, device=None) -> float:
    """`mse_with_mask_loss` divides by the PADDED size bs * Nmax (src/utils/utils.py:55).  For the
    data-parallel step to equal the single-process step on the concatenated batch the denominator
    must be (sum of bs over ranks) * (max Nmax over ranks)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(local_batch * local_nmax)
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    bs = torch.tensor([local_batch], dtype=torch.int64, device=dev)
    nm = torch.tensor([local_nmax], dtype=torch.int64, device=dev)
    dist.all_reduce(bs, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(nm, op=dist.ReduceOp.MAX, group=group)
    return float(int(bs.item()) * int(nm.item()))


def allreduce_gradients(params, group=None, average: bool = False) -> None:
    """One flat all-reduce (sum) over every gradient: 3.4 M fp32 values = 13.7 MB for the benchmark
    model, i.e. latency-bound on NVLink -- a single bucket is the right size.  With the loss
    normalised by `global_loss_denominator` the sum IS the gradient of the global batch; pass
    `average=True` when every rank normalised by its own local size instead."""
    params = [p for p in params if p.grad is not None]
    if not params or not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    # The native backward hands out views of ONE flat buffer: when every gradient still lives in that storage
    # (and fills it), reduce the storage in place -- no concatenation, no copy back.
    st = params[0].grad.untyped_storage()
    total = sum(p.grad.numel() for p in params)
    if (all(p.grad.dtype == torch.float32 and p.grad.is_contiguous() and p.grad.untyped_storage().data_ptr() == st.data_ptr()
            for p in params) and st.nbytes() == 4 * total
            and len({p.grad.data_ptr() for p in params}) == len(params)):
        base = torch.empty(0, dtype=torch.float32, device=params[0].grad.device).set_(st, 0, (total,))
        dist.all_reduce(base, op=dist.ReduceOp.SUM, group=group)
        if average:
            base.div_(world)
        return
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    off = 0
    for p in params:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


# ---------------------------------------------------------------------------------------------
# data-parallel training step that overlaps: per-layer gradient buckets, no host round trip
# ---------------------------------------------------------------------------------------------
def bucket_slices(embed_n: int, layer_n: int, num_layers: int, total: int):
    """Slices of the flat gradient buffer ([embedding | layer 0 | ... | layer L-1 | head], the layout
    `_ScorerTrainFn.backward` hands out) in the order the backward completes them: bucket L-1 (with the head's
    two tensors, adjacent in memory and finished first), L-2, ..., 0, then -1 = the embedding."""
    out = {}
    for l in range(num_layers - 1, -1, -1):
        lo = embed_n + l * layer_n
        out[l] = (lo, total if l == num_layers - 1 else lo + layer_n)
    out[-1] = (0, embed_n)
    return out


def dp_extras(loss_sum, batch: int, nmax: int, rank: int, world: int) -> torch.Tensor:
    """[sum of squared errors, batch size, Nmax one-hot over ranks]: after a SUM all-reduce every rank holds the global
    sum, the global batch size and every rank's Nmax, so the padded-size denominator of `mse_with_mask_loss`
    (src/utils/utils.py:55) for the GLOBAL batch comes out of the same kind of collective as the gradients."""
    ext = torch.zeros(2 + world, dtype=torch.float32)
    ext[0], ext[1], ext[2 + rank] = float(loss_sum), float(batch), float(nmax)
    return ext


def dp_denominator(ext_reduced: torch.Tensor) -> float:
    """(sum of batch sizes) * (max Nmax) from SUM-reduced `dp_extras` (what vsum_dp_finalize computes on the device)."""
    return float(ext_reduced[1]) * float(ext_reduced[2:].max())


class DataParallel:
    """Data-parallel wrapper of a `SimNet` for the training step of `src/train.py:111-131` on several GPUs (one process
    per GPU, NCCL).  Equal to the single-process step on the concatenated batch, like `allreduce_gradients` +
    `global_loss_denominator`, but

    * the gradient all-reduce is issued PER LAYER from a communication stream as `vsum_scorer_backward_hooked` finishes
      queuing each layer (the last layer's bucket also carries the head), so only the embedding bucket is exposed; steps
      too small to hide a collective (fewer than `bucket_min_frames` frames on this rank: the step is launch-bound) use
      ONE all-reduce after the backward instead;
    * nothing synchronises the host: every rank back-propagates the UN-normalised sum of squared errors, the global
      denominator (sum of batch sizes) x (max Nmax) is derived on the device from extras that ride the first bucket's SUM
      all-reduce (they sit right behind the head's gradients in the flat buffer), and `vsum_dp_finalize` scales the
      reduced gradients and the loss.

        ddp = DataParallel(model)
        out, _ = model(x, mask); ddp.loss(out, targets, mask).backward(); loss = ddp.finish(); optimizer.step()
    """

    def __init__(self, model, group=None, bucket_min_frames: int = 8192):
        from . import _cabi
        self.model, self.group = model, group
        self.active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.active else 1
        self.rank = dist.get_rank(group) if self.active else 0
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise _cabi.VsumError("DataParallel needs the model on a CUDA device (no CPU fallback)")
        self.bucket_min_frames = int(bucket_min_frames)
        self.ext_n = 2 + self.world                                # floats appended to the flat gradient buffer
        self._comm = torch.cuda.Stream(self.device)
        self._ext = torch.zeros(self.ext_n, dtype=torch.float32, device=self.device)
        self._loss_out = torch.zeros((), dtype=torch.float32, device=self.device)
        self._flat = None
        self._works, self._error = [], None
        model._dp = self

    def detach(self):
        self.model._dp = None

    def loss(self, output, targets, mask, batch: int = None, nmax: int = None):
        """Sum of squared errors of this rank's valid frames (what `.backward()` starts from); `finish()` returns the
        global masked MSE.  `mask` bool [bs, Nmax], True = padded frame; callers that hold packed rows instead of a padded
        batch pass `batch` (videos) and `nmax` (longest video) explicitly."""
        from . import _cabi
        from .utils import mse_with_mask_loss
        s = mse_with_mask_loss(output, targets, mask, denom=1.0)
        with torch.cuda.device(self.device):
            _cabi.check(_cabi.load().vsum_dp_extras(self._ext.data_ptr(), s.detach().data_ptr(),
                                                    int(mask.shape[0] if batch is None else batch),
                                                    int(mask.shape[1] if nmax is None else nmax), self.rank, self.world,
                                                    torch.cuda.current_stream(self.device).cuda_stream), "vsum_dp_extras")
        return s

    # ---- called from _ScorerTrainFn.backward -------------------------------------------------
    def _begin(self, flat, n_grads, embed_n, layer_n, num_layers, frames):
        """`flat` = [gradients (n_grads) | extras (ext_n)], zeroed by the caller."""
        self._flat, self._n_grads = flat, n_grads
        self._slices = bucket_slices(embed_n, layer_n, num_layers, flat.numel())     # the last layer's bucket runs to the end: head + extras
        self._bucketed = frames >= self.bucket_min_frames
        self._works, self._error = [], None
        flat[n_grads:].copy_(self._ext)

    def _reduce(self, tensor):
        if self.world == 1:
            return
        main = torch.cuda.current_stream(self.device)
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(ev)
            self._works.append(dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def _bucket_ready(self, bucket: int):
        try:                                                        # (an exception must not unwind through the C caller)
            if self._bucketed:
                lo, hi = self._slices[bucket]
                self._reduce(self._flat[lo:hi])
            elif bucket == -1 and self.world > 1:
                # launch-bound step: nothing to overlap with -- one all-reduce on the compute stream itself, no events, no
                # second stream (the host, not the GPU, is what such a step waits for)
                dist.all_reduce(self._flat, op=dist.ReduceOp.SUM, group=self.group)
        except BaseException as e:
            self._error = e

    def finish(self) -> torch.Tensor:
        """Queues the end of the step: wait for the all-reduces on the communication stream, scale gradients and loss by
        the global denominator there, make the current stream wait.  Returns the global loss (device scalar)."""
        from . import _cabi
        if self._error is not None:
            raise self._error
        if self._flat is None:
            raise _cabi.VsumError("DataParallel.finish: no backward has run since the last step")
        flat, self._flat = self._flat, None
        main = torch.cuda.current_stream(self.device)
        plist = self.model._param_list() if hasattr(self.model, "_param_list") else list(self.model.parameters())
        params = [p for p in plist if p.grad is not None]
        st = flat.untyped_storage().data_ptr()
        if not all(p.grad.untyped_storage().data_ptr() == st for p in params):
            raise _cabi.VsumError("DataParallel.finish: a .grad no longer lives in the backward's flat buffer "
                                  "(use optimizer.zero_grad(set_to_none=True) and one backward per step)")
        if not self._bucketed:                                      # everything already sits on the compute stream
            _cabi.check(_cabi.load().vsum_dp_finalize(flat.data_ptr(), self._n_grads, flat.data_ptr() + 4 * self._n_grads, self.world,
                                                      self._loss_out.data_ptr(), main.cuda_stream), "vsum_dp_finalize")
            return self._loss_out
        with torch.cuda.stream(self._comm):
            self._comm.wait_stream(main)
            for w in self._works:
                w.wait()
            _cabi.check(_cabi.load().vsum_dp_finalize(flat.data_ptr(), self._n_grads, flat.data_ptr() + 4 * self._n_grads, self.world,
                                                      self._loss_out.data_ptr(), self._comm.cuda_stream), "vsum_dp_finalize")
        self._works = []
        main.wait_stream(self._comm)
        return self._loss_out
