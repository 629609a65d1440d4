"""CPU: the only exchange step of the multi-GPU path (all-gather of per-video F-scores) on a
2-rank gloo group, plus shard-independence of the final mean."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from vsum_b200.sharding import gather_fscores, partition


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ns, f_all, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = partition(ns, world)[rank]
    out = gather_fscores(mine, f_all[mine], len(ns))
    ret[rank] = out.tobytes()
    dist.barrier()
    dist.destroy_process_group()


def test_gather_fscores_two_ranks():
    rng = np.random.default_rng(0)
    ns = [int(x) for x in rng.integers(128, 4000, 37)]
    f_all = rng.random(37) * 100
    f_all[5] = np.nan                                   # NaN F-scores (empty summary) must survive the gather
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ns, f_all, ret), nprocs=2, join=True)
    for rank in (0, 1):
        got = np.frombuffer(ret[rank], dtype=np.float64)
        assert got.tobytes() == f_all.tobytes()          # bit-identical, video order, on every rank


def test_single_process_path():
    f = gather_fscores([2, 0], np.array([5.0, 7.0]), 3)
    assert f[0] == 7.0 and f[2] == 5.0 and np.isnan(f[1])


def _grad_worker(rank, world, port, ret):
    import torch
    from vsum_b200.sharding import allreduce_gradients, global_loss_denominator
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Linear(6, 3)
    for i, p in enumerate(lin.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    allreduce_gradients(lin.parameters())
    denom = global_loss_denominator(local_batch=4 - rank, local_nmax=100 + 50 * rank)
    ret[rank] = ([p.grad.flatten()[0].item() for p in lin.parameters()], denom)
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_allreduce_and_global_denominator_two_ranks():
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_grad_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        grads, denom = ret[rank]
        assert grads == [3.0, 6.0]                      # (1 + 2) * (i + 1): summed over the two ranks
        assert denom == float((4 + 3) * 150)            # sum of batch sizes x max padded length
