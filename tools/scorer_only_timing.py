"""GPU box: the scorer alone on the bench workload (no evaluation stream) -- how much of the pipelined step is interference?"""
import sys, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200.model import SimNet
from vsum_b200.synthetic import video_length
lens = [video_length(v, 128, 8192) for v in range(256)]
T = sum(lens)
torch.manual_seed(1234)
model = SimNet(num_heads=4, d_model=256, num_layers=4, sparsity=0., use_cls=False, dropout=0.3, num_classes=1, use_pos=True).cuda().eval()
x = torch.rand((T, 1024), device="cuda")
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
out = torch.empty((T, 1), device="cuda")
for _ in range(3): model.forward_packed(x, cu, lens, apply_sigmoid=True, want_feats=False, scores_out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): model.forward_packed(x, cu, lens, apply_sigmoid=True, want_feats=False, scores_out=out)
e1.record(); torch.cuda.synchronize()
print(f"scorer only: {e0.elapsed_time(e1) / 10:.3f} ms per 256 videos ({T} frames)")
