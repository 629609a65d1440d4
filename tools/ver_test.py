import sys, torch, numpy as np
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
from vsum_b200.model import SimNet
from vsum_b200.utils import mse_with_mask_loss
torch.manual_seed(0)
for fused in (True, False):
    model = SimNet(num_heads=4, d_model=256, num_layers=4, sparsity=0., dropout=0.0).cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=fused)
    T = 600
    feats = torch.randn((T, 1024), device="cuda"); tgt = torch.rand((1, T), device="cuda")
    cu = torch.tensor([0, 300, 600], dtype=torch.int32, device="cuda"); nopad = torch.zeros((1, T), dtype=torch.bool, device="cuda")
    p0 = next(model.parameters())
    for it in range(6):
        n0 = _cabi.launch_count(); v0 = p0._version
        opt.zero_grad(set_to_none=True)
        out, _ = model.forward_packed_train(feats, cu, [300, 300])
        loss = mse_with_mask_loss(out.view(1, T, 1), tgt, nopad)
        loss.backward(); opt.step()
        print("fused", fused, "step", it, "loss", float(loss), "launches", _cabi.launch_count() - n0, "version", v0, "->", p0._version)
