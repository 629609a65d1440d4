"""GPU box: how far are q/k weight gradients of torch's own bf16 autocast from the fp32 gradients, on the batch
tests/test_train_gpu.py uses?  Yardstick for the bf16 training mode's tolerance."""
import sys, torch
sys.path.insert(0, "video-summarization_b200"); sys.path.insert(0, "."); sys.path.insert(0, "tests")
from oracle import scorer_ref
from vsum_b200.model import SimNet
from vsum_b200.utils import mse_with_mask_loss
from test_train_gpu import padded_batch

torch.manual_seed(11)
kw = dict(num_heads=4, d_model=256, num_layers=4, dropout=0.0)
model = SimNet(sparsity=0., use_cls=False, num_classes=1, use_pos=True, **kw).cuda()
with torch.no_grad():
    for p in model.parameters():
        p.add_(0.05 * torch.randn_like(p))
model.eval()
x, tgt, mask = padded_batch((300, 513))
x, tgt, mask = x.cuda(), tgt.cuda(), mask.cuda()
def grads(autocast):
    params = {k: v.detach().cpu().clone().requires_grad_(v.dtype.is_floating_point and "pos_embedding" not in k) for k, v in model.state_dict().items()}
    with torch.enable_grad(), torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast):
        logits, _ = scorer_ref.scorer_forward.__wrapped__(params, x.cpu(), 4, mask.cpu())
        loss = scorer_ref.masked_mse(logits.float(), tgt.cpu(), mask.cpu())
    loss.backward()
    return {k: p.grad.cuda() for k, p in params.items() if p.requires_grad}
ref, amp = grads(False), grads(True)
res = {}
for prec in ("tf32", "bf16"):
    model.zero_grad(); model.train_precision = prec
    pred, _ = model(x, mask)
    mse_with_mask_loss(pred, tgt, mask).backward()
    res[prec] = {k: p.grad.clone() for k, p in model.named_parameters()}
print(f"{'parameter':50s} {'torch-amp-bf16':>14s} {'vsum tf32':>10s} {'vsum bf16':>10s}   (relative Frobenius error vs fp32 autograd)")
for k, g in ref.items():
    if ".sa." in k and "weight" in k or "fc1.weight" in k or "feature_transform.weight" in k:
        r = lambda o: ((o - g).norm() / g.norm()).item()
        print(f"{k:50s} {r(amp[k]):14.3e} {r(res['tf32'][k]):10.3e} {r(res['bf16'][k]):10.3e}")
