"""GPU (needs >= 2 devices, skipped otherwise): the data-parallel finetune step of BASELINE config 3.
Two ranks, one batch shard each, NCCL all-reduce of the flat gradient: the result must equal the
single-process step on the concatenated batch (same loss denominator bs_total * Nmax_global)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

LENS = [(140, 90), (60, 200)]          # rank 0 / rank 1 shards


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _batch(lens, first_id):
    from vsum_b200.synthetic import make_video
    nmax = max(lens)
    x = torch.full((len(lens), nmax, 1024), 1000.0)
    t = torch.full((len(lens), nmax), 1000.0)
    for b, n in enumerate(lens):
        v = make_video(first_id + b, n)
        x[b, :n] = torch.from_numpy(v.features)
        t[b, :n] = torch.from_numpy(v.gtscore)
    return x, t


def _model():
    from vsum_b200.model import SimNet
    torch.manual_seed(21)
    m = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.0)
    m.train_precision = "fp32"      # the overlapped path scales AFTER the backward: exact in fp32 up to rounding, a different bf16 rounding otherwise
    return m


def _worker(rank, world, port, ret, overlapped=False):
    from vsum_b200.sharding import DataParallel, allreduce_gradients, global_loss_denominator
    from vsum_b200.utils import mse_with_mask_loss
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    model = _model().cuda().train()
    x, t = _batch(LENS[rank], 1300 + 10 * rank)
    x, t = x.cuda(), t.cuda()
    mask = x[:, :, 0] == 1000
    if overlapped:          # per-layer buckets from a communication stream, denominator derived on the device
        ddp = DataParallel(model, bucket_min_frames=0)        # per-layer buckets even for this small step
        pred, _ = model(x, mask)
        ddp.loss(pred, t, mask).backward()
        ret[f"loss{rank}"] = float(ddp.finish())
    else:
        denom = global_loss_denominator(x.shape[0], x.shape[1])
        pred, _ = model(x, mask)
        loss = mse_with_mask_loss(pred, t, mask, denom=denom)
        loss.backward()
        allreduce_gradients(model.parameters())
    torch.cuda.synchronize()
    ret[rank] = {k: p.grad.cpu().numpy() for k, p in model.named_parameters()}
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlapped", [False, True], ids=["flat_allreduce", "bucketed_overlapped"])
def test_two_rank_step_equals_single_process_step(overlapped):
    from vsum_b200.utils import mse_with_mask_loss
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, ret, overlapped), nprocs=2, join=True)
    # single process, concatenated batch padded to the global maximum
    model = _model().cuda().train()
    lens = LENS[0] + LENS[1]
    nmax = max(lens)
    xs, ts = [], []
    for r in range(2):
        x, t = _batch(LENS[r], 1300 + 10 * r)
        pad = nmax - x.shape[1]
        xs.append(torch.nn.functional.pad(x, (0, 0, 0, pad), value=1000.0))
        ts.append(torch.nn.functional.pad(t, (0, pad), value=1000.0))
    x, t = torch.cat(xs).cuda(), torch.cat(ts).cuda()
    mask = x[:, :, 0] == 1000
    pred, _ = model(x, mask)
    single = mse_with_mask_loss(pred, t, mask)
    single.backward()
    if overlapped:
        for rank in (0, 1):
            assert abs(ret[f"loss{rank}"] - float(single)) <= 1e-5 * abs(float(single))
    for k, p in model.named_parameters():
        want = p.grad.cpu().numpy()
        for rank in (0, 1):
            np.testing.assert_allclose(ret[rank][k], want, rtol=2e-4, atol=1e-7 + 2e-5 * np.abs(want).max())
