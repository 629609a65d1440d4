"""GPU box: time one training step (forward with tape + masked MSE + backward) of the scorer at both
train_precision settings and list the kernels by time (torch.profiler / CUPTI sees the ctypes launches)."""
import sys, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200.model import SimNet
from vsum_b200.utils import mse_with_mask_loss

lens = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [320, 280, 410, 350]   # TVSum-like batch of 4
torch.manual_seed(0)
T, B = sum(lens), len(lens)
feats = torch.randn(T, 1024, device="cuda")
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
tgt = torch.rand(1, T, device="cuda")
nopad = torch.zeros(1, T, dtype=torch.bool, device="cuda")
for prec in ("fp32", "tf32", "bf16"):
    model = SimNet(num_heads=4, d_model=256, num_layers=4, sparsity=0., dropout=float(__import__("os").environ.get("DROP", "0.3"))).cuda().train()
    model.train_precision = prec
    def step():
        model.zero_grad(set_to_none=True)
        out, _ = model.forward_packed_train(feats, cu, lens)
        loss = mse_with_mask_loss(out.view(1, T, 1), tgt, nopad)
        loss.backward()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): step()
    e1.record(); torch.cuda.synchronize()
    print(f"train_precision={prec} T={T} B={B}: {e0.elapsed_time(e1)/10:.3f} ms/step", flush=True)
    if "--kernels" in sys.argv:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as pr:
            step(); torch.cuda.synchronize()
        print(pr.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70), flush=True)
