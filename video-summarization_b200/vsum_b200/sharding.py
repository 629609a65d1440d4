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
