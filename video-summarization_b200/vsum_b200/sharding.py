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
