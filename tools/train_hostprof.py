"""GPU box: where the HOST time of a finetune step goes (cProfile over 200 steps, no synchronisation inside the loop)."""
import cProfile, pstats, sys, time, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200.model import SimNet
from vsum_b200.utils import mse_with_mask_loss
torch.manual_seed(1234)
model = SimNet(num_heads=4, d_model=256, num_layers=4, sparsity=0., dropout=0.3).cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, fused=True)
lens = [311, 257, 380, 222]; T = sum(lens)
feats = torch.randn((T, 1024), device="cuda"); tgt = torch.rand((1, T), device="cuda")
nopad = torch.zeros((1, T), dtype=torch.bool, device="cuda")
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
def step():
    opt.zero_grad(set_to_none=True)
    out, _ = model.forward_packed_train(feats, cu, lens)
    loss = mse_with_mask_loss(out.view(1, T, 1), tgt, nopad)
    loss.backward()
    opt.step()
for _ in range(20): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): step()
t_host = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print(f"host issue time {t_host / 200 * 1e3:.3f} ms/step, wall {t_all / 200 * 1e3:.3f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(200): step()
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
