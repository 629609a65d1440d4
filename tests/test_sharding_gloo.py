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


def _extras_worker(rank, world, port, ret):
    import torch
    from vsum_b200.sharding import dp_denominator, dp_extras, global_loss_denominator
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    bs, nmax, loss_sum = 4 - rank, 100 + 50 * rank, 2.5 * (rank + 1)
    ext = dp_extras(loss_sum, bs, nmax, rank, world)
    dist.all_reduce(ext, op=dist.ReduceOp.SUM)
    ret[rank] = (dp_denominator(ext), float(ext[0]), global_loss_denominator(bs, nmax))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_extras_give_the_global_denominator_from_one_sum_allreduce():
    """DataParallel's host-sync-free normalisation: batch size and a one-hot Nmax ride a SUM all-reduce; the denominator
    derived from them equals global_loss_denominator (two blocking collectives) on every rank."""
    from vsum_b200.sharding import bucket_slices
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_extras_worker, args=(2, port, ret), nprocs=2, join=True)
    for rank in (0, 1):
        d, s, want = ret[rank]
        assert d == want == float(7 * 150) and s == 7.5
    sl = bucket_slices(embed_n=10, layer_n=7, num_layers=3, total=10 + 21 + 4)
    assert sl == {2: (24, 35), 1: (17, 24), 0: (10, 17), -1: (0, 10)}        # the last layer's bucket carries the head (and the extras)
    covered = sorted(sl.values())
    assert covered[0][0] == 0 and covered[-1][1] == 35 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))
