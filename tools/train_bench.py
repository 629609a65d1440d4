"""Training-step throughput of the scorer on B200.

  python tools/train_bench.py --config finetune            # bs 4 / GPU, N ~ U[200,400]   (run_finetune.sh:1)
  python tools/train_bench.py --config pretrain --len 2048 # bs 8 / GPU of N-frame videos (run_pretrain.sh:1)
  python -m torch.distributed.run --nproc-per-node 8 ... tools/train_bench.py --config finetune

One step = forward with tape + masked MSE + backward through the C-ABI (+ one flat NCCL all-reduce of the
gradients when world > 1) + Adam.  Features are resident on the device; timing is CUDA events, max over ranks.
Prints one JSON line (frames/s and videos/s over all ranks, algorithmic training TFLOP/s = 3 x forward)."""
import argparse, json, os, sys
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "video-summarization_b200"))
from vsum_b200 import _cabi
from vsum_b200.model import SimNet
from vsum_b200.sharding import allreduce_gradients, global_loss_denominator, scorer_cost
from vsum_b200.utils import mse_with_mask_loss


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="finetune", choices=["finetune", "pretrain"])
    ap.add_argument("--len", type=int, default=2048)
    ap.add_argument("--batch", type=int, default=0, help="videos per GPU per step (default 4 finetune / 8 pretrain)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    args = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    bs = args.batch or (4 if args.config == "finetune" else 8)
    rng = np.random.default_rng(99 + rank)
    lens = [int(n) for n in (rng.integers(200, 401, bs) if args.config == "finetune" else [args.len] * bs)]
    T = sum(lens)
    torch.manual_seed(1234)
    model = SimNet(num_heads=4, d_model=256, num_layers=4, sparsity=0., dropout=0.3).cuda().train()
    model.train_precision = args.precision
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
    g = torch.Generator(device="cuda").manual_seed(7 + rank)
    feats = torch.randn((T, 1024), device="cuda", generator=g)
    tgt = torch.rand((1, T), device="cuda", generator=g)
    nopad = torch.zeros((1, T), dtype=torch.bool, device="cuda")
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    denom = global_loss_denominator(bs, max(lens))          # utils.py:55 semantics for the global batch

    def step():
        opt.zero_grad(set_to_none=True)
        out, _ = model.forward_packed_train(feats, cu, lens)
        loss = mse_with_mask_loss(out.view(1, T, 1), tgt, nopad, denom=denom)
        loss.backward()
        allreduce_gradients(model.parameters())
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = _cabi.load().vsum_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    launches = (_cabi.load().vsum_launch_count() - n0) // args.steps
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps, float(T), float(bs), sum(scorer_cost(n) for n in lens)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = ms.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms, op=dist.ReduceOp.SUM)
        ms[0] = mx[0]
    t_ms, frames, videos, fwd_flops = (float(v) for v in ms.tolist())
    if rank == 0:
        print(json.dumps({
            "metric": "train_frames_per_sec", "value": frames / t_ms * 1e3, "unit": "frames/s", "videos_per_sec": videos / t_ms * 1e3,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms, "higher_is_better": True,
            "scaling": "weak", "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"{args.config}: {bs} videos/GPU/step, lens {lens if bs <= 8 else lens[:8]} (rank 0), d256/h4/L4, dropout 0.3, "
                                   "forward+tape, masked MSE, backward, flat gradient all-reduce, fused Adam"},
            "train_tflops": 3.0 * fwd_flops / t_ms / 1e9, "gpu_launches_per_step": int(launches), "final_loss": float(loss.item())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
