"""GPU parity of the native training path (fp32 kernels): loss and EVERY parameter gradient of
SimNet + masked MSE against PyTorch autograd through the fp32 CPU oracle (oracle/scorer_ref.py),
on padded batches exactly as src/train.py:111-131 builds them."""
import numpy as np
import pytest
import torch

from oracle import scorer_ref
from vsum_b200.model import SimNet
from vsum_b200.synthetic import make_video
from vsum_b200.utils import mse_with_mask_loss

pytestmark = pytest.mark.gpu


def padded_batch(lens, first_id=1200):
    nmax = max(lens)
    x = torch.full((len(lens), nmax, 1024), 1000.0)
    tgt = torch.full((len(lens), nmax), 1000.0)
    for b, n in enumerate(lens):
        v = make_video(first_id + b, n)
        x[b, :n] = torch.from_numpy(v.features)
        tgt[b, :n] = torch.from_numpy(v.gtscore)
    return x, tgt, x[:, :, 0] == 1000                       # collate_fn_train + train.py:118


def oracle_grads(sd, x, tgt, mask, num_heads):
    params = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "pos_embedding" not in k) for k, v in sd.items()}
    with torch.enable_grad():
        logits, feats = scorer_ref.scorer_forward.__wrapped__(params, x, num_heads, mask)
        loss = scorer_ref.masked_mse(logits, tgt, mask)
    loss.backward()
    return loss.item(), {k: p.grad for k, p in params.items() if p.requires_grad}


@pytest.mark.parametrize("kw,lens,train_precision,tol", [
    (dict(num_heads=4, d_model=256, num_layers=2, dropout=0.3), (150, 97, 64, 130), "fp32", 2e-4),
    (dict(num_heads=4, d_model=64, num_layers=3, dropout=0.1), (70, 33), "fp32", 2e-4),
    (dict(num_heads=4, d_model=256, num_layers=4, dropout=0.0), (300,), "fp32", 2e-4),
    # linear layers on the tensor cores: tf32 forward / dgrad, bf16 wgrad (8-bit mantissa operands)
    (dict(num_heads=4, d_model=256, num_layers=2, dropout=0.3), (150, 97, 64, 130), "tf32", 2e-2),
    (dict(num_heads=4, d_model=256, num_layers=4, dropout=0.0), (300,), "tf32", 2e-2),
    # + attention forward / backward on the tensor cores with bf16 operands (delta is taken from the same
    # rounded dO the MMAs see, so dS rows sum to zero and the small q/k gradients stay accurate:
    # profiles/r01_train_grad_error_yardstick.txt, 2x closer to fp32 than torch's own bf16 autocast)
    (dict(num_heads=4, d_model=256, num_layers=2, dropout=0.3), (150, 97, 64, 130), "bf16", 3e-2),
    (dict(num_heads=4, d_model=256, num_layers=4, dropout=0.0), (300, 513), "bf16", 3e-2),
])
def test_gradients_match_autograd_oracle(kw, lens, train_precision, tol):
    torch.manual_seed(11)
    model = SimNet(sparsity=0., use_cls=False, num_classes=1, use_pos=True, **kw).cuda()
    with torch.no_grad():
        for p in model.parameters():                        # non-trivial LayerNorm affine / biases
            p.add_(0.05 * torch.randn_like(p))
    model.eval()                                            # dropout off: deterministic comparison
    model.train_precision = train_precision
    x, tgt, mask = padded_batch(lens)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    want_loss, want = oracle_grads(sd, x, tgt, mask, kw["num_heads"])

    pred, feats = model(x.cuda(), mask.cuda())
    assert pred.requires_grad and pred.shape == (len(lens), max(lens), 1)
    loss = mse_with_mask_loss(pred, tgt.cuda(), mask.cuda())
    loss.backward()
    assert abs(loss.item() - want_loss) <= (1e-5 if train_precision == "fp32" else 6e-3) * max(1.0, abs(want_loss))
    named = dict(model.named_parameters())
    assert set(named) == set(want)
    # k.bias has a mathematically zero gradient (softmax is invariant to it): in the reduced-precision modes
    # its rounding noise is judged against the q/k/v bias gradients' scale instead of against ~0
    floor = 1e-6 if train_precision == "fp32" else 1e-2 * max(g.abs().max().item() for k, g in want.items() if ".sa." in k and "bias" in k)
    for k, g in want.items():
        got = named[k].grad
        assert got is not None, k
        scale = max(g.abs().max().item(), floor)
        err = (got.cpu() - g).abs().max().item()
        assert err <= tol * scale + 1e-7, f"{k}: max err {err:.3e} vs scale {scale:.3e}"
        if train_precision != "fp32" and g.abs().max().item() > floor:
            rel = ((got.cpu() - g).norm() / g.norm()).item()
            assert rel <= 1e-2, f"{k}: relative Frobenius error {rel:.3e}"


def test_dropout_is_active_and_consistent():
    torch.manual_seed(3)
    model = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.3).cuda().train()
    model.train_precision = "fp32"                          # the finite-difference check below needs fp32 linears
    x, tgt, mask = padded_batch((120, 60))
    x, tgt, mask = x.cuda(), tgt.cuda(), mask.cuda()
    torch.manual_seed(5)
    a, _ = model(x, mask)
    torch.manual_seed(5)
    b, _ = model(x, mask)
    torch.manual_seed(6)
    c, _ = model(x, mask)
    assert torch.equal(a, b)                                # same seed -> same masks
    assert not torch.equal(a, c)                            # different seed -> different masks
    model.eval()
    with torch.no_grad():
        e, _ = model(x, mask)
    assert not torch.allclose(a, e)
    # finite-difference check of the dropout backward along a random direction (fixed masks)
    model.train()
    p = model.encoder.module_list[0].mlp.fc1.weight
    direction = torch.randn_like(p)
    def loss_at(eps):
        with torch.no_grad():
            p.add_(eps * direction)
        torch.manual_seed(9)
        pred, _ = model(x, mask)
        val = mse_with_mask_loss(pred, tgt, mask)
        with torch.no_grad():
            p.sub_(eps * direction)
        return val
    torch.manual_seed(9)
    pred, _ = model(x, mask)
    model.zero_grad()
    mse_with_mask_loss(pred, tgt, mask).backward()
    analytic = (p.grad * direction).sum().item()
    h = 3e-3                                                # small enough that few ReLU / LayerNorm kinks are crossed
    numeric = (loss_at(h).item() - loss_at(-h).item()) / (2 * h)
    assert abs(analytic - numeric) <= 5e-2 * max(abs(numeric), 1e-3), (analytic, numeric)


def test_adam_step_reduces_loss():
    """train.py:124-127 with GradScaler: a few optimiser steps on one batch must reduce the loss."""
    torch.manual_seed(0)
    model = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.0).cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    scaler = torch.amp.GradScaler("cuda")
    x, tgt, mask = padded_batch((100, 80, 40))
    x, tgt, mask = x.cuda(), tgt.cuda(), mask.cuda()
    losses = []
    for _ in range(12):
        with torch.autocast("cuda"):
            pred, _ = model(x, mask)
            loss = mse_with_mask_loss(pred, tgt, mask)
        opt.zero_grad()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        losses.append(loss.item())
    assert max(losses[-3:]) < losses[0] * 0.9, losses


def test_pretrain_losses_match_reference_and_backprop():
    """PretrainModel (src/model/simnet_pretrain.py:71-100): the three losses against the reference's
    own values (golden fixture), and gradients reach the encoder through the native backward."""
    import os
    from conftest import GOLDEN
    from vsum_b200.model import PretrainModel
    g = np.load(os.path.join(GOLDEN, "pretrain_golden.npz"))["losses"]
    torch.manual_seed(1234)
    import random; random.seed(1234); np.random.seed(1234)
    net = PretrainModel(feature_dim=256, sparsity=0.0, num_heads=4, num_layers=4, dropout=0.2, use_pos=True).cuda().eval()
    net.encoder.train_precision = "fp32"                     # exact mode for the comparison with the reference values
    lens = (120, 77)
    x = torch.full((2, 120, 1024), 1000.0)
    for b, n in enumerate(lens):
        x[b, :n] = torch.from_numpy(make_video(130 + b, n).features)
    vid_rep = torch.from_numpy(np.random.default_rng(4321).random((2, 512), dtype=np.float32)).cuda()
    x = x.cuda()
    mask = x[:, :, 0] == 1000
    loss, center, repel = net(x, vid_rep, mask)
    np.testing.assert_allclose([loss.item(), center.item(), repel.item()], g, rtol=2e-4, atol=1e-6)
    (loss + 0.5 * center + 1.0 * repel).backward()                      # pretrain.py:63
    grads = [p.grad for p in net.encoder.parameters()]
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads)
    assert sum(gr.abs().sum().item() for gr in grads) > 0


def _reference_losses(scores, x512, mask, video_rep, t, pen_met):
    """torch restatement of simnet_pretrain.py:35-100 on padded tensors, [N,N] matrix and all (the checker)."""
    import torch.nn.functional as F
    n = x512.shape[1]
    xm = x512 * (~mask).unsqueeze(2)
    xh = xm / (xm.norm(dim=2, keepdim=True) + 1e-9)
    sim = torch.matmul(xh, xh.transpose(1, 2)) * (torch.eye(n, device=x512.device) == 0).float().unsqueeze(0)
    repel = sim.mean(dim=1).mean()
    m3 = mask.unsqueeze(2)
    mix = F.softmax(scores.masked_fill(m3, float("-inf")) / t, dim=1)
    if pen_met == "entropy":
        e = (mix + 1e-9) * torch.log(mix + 1e-9)
        center = e.masked_fill(m3, 0.).mean(dim=1).mean()
    else:
        center = torch.norm(mix, dim=1).mean()
    pooled = torch.matmul(mix.transpose(1, 2), x512).squeeze(1)
    loss = (-F.softmax(video_rep, dim=1) * torch.log(F.softmax(pooled, dim=1))).mean()
    return loss, center, repel


@pytest.mark.parametrize("lens,pen_met", [((120, 77), "entropy"), ((300, 513, 1), "entropy"), ((64, 200), "norm")])
def test_pretrain_loss_kernels_forward_backward(lens, pen_met):
    """vsum_pretrain_losses_* and vsum_linear_* (fp32 mode) against autograd through the restated reference formulas."""
    from vsum_b200.model.simnet_pretrain import _LinearFn, _PretrainLossFn
    torch.manual_seed(17)
    B, n = len(lens), max(lens)
    T = sum(lens)
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    feats = torch.randn(T, 256, device="cuda", dtype=torch.float64)
    scores = torch.randn(T, 1, device="cuda", dtype=torch.float64)
    w = (torch.randn(512, 256, device="cuda", dtype=torch.float64) / 16)
    b = torch.randn(512, device="cuda", dtype=torch.float64) * 0.1
    vid_rep = torch.rand(B, 512, device="cuda")
    coef = torch.tensor([1.0, 0.5, 1.0], device="cuda")                    # pretrain.py:63
    # checker: fp64 autograd on the padded layout
    ref_in = [v.clone().requires_grad_(True) for v in (feats, scores, w, b)]
    pf = torch.zeros(B, n, 256, device="cuda", dtype=torch.float64)
    ps = torch.zeros(B, n, 1, device="cuda", dtype=torch.float64)
    mask = torch.ones(B, n, dtype=torch.bool, device="cuda")
    for v, ln in enumerate(lens):
        mask[v, :ln] = False
    idx = (~mask).reshape(-1).nonzero().squeeze(1)
    pf = pf.reshape(B * n, 256).index_copy(0, idx, ref_in[0]).reshape(B, n, 256)
    ps = ps.reshape(B * n, 1).index_copy(0, idx, ref_in[1]).reshape(B, n, 1)
    want = torch.stack(_reference_losses(ps, pf @ ref_in[2].t() + ref_in[3], mask, vid_rep.double(), 0.4, pen_met))
    (want * coef.double()).sum().backward()
    # native kernels, fp32
    got_in = [v.float().clone().requires_grad_(True) for v in (feats, scores, w, b)]
    x512 = _LinearFn.apply(got_in[0], got_in[2], got_in[3], 0)
    got = _PretrainLossFn.apply(got_in[1], x512, cu, list(lens), n, 0.4, vid_rep, pen_met == "entropy")
    (got * coef).sum().backward()
    np.testing.assert_allclose(got.detach().cpu().numpy(), want.detach().cpu().numpy(), rtol=2e-5, atol=1e-7)
    for name, a, r in zip(("d_feats", "d_scores", "d_w", "d_b"), got_in, ref_in):
        scale = r.grad.abs().max().item()
        err = (a.grad.double() - r.grad).abs().max().item()
        assert err <= 2e-4 * scale + 1e-10, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
    # tensor-core mode of the linear layer
    tc_in = [v.float().clone().requires_grad_(True) for v in (feats, w, b)]
    y = _LinearFn.apply(tc_in[0], tc_in[1], tc_in[2], 1)
    gy = torch.randn_like(y)
    y.backward(gy)
    ref_y = feats @ w.t() + b
    assert (y.double() - ref_y).abs().max().item() <= 5e-3 * ref_y.abs().max().item()
    for a, r in zip(tc_in, (gy.double() @ w, gy.double().t() @ feats, gy.double().sum(0))):
        assert (a.grad.double() - r).abs().max().item() <= 2e-2 * r.abs().max().item()


def test_data_parallel_wrapper_single_rank_equals_plain_step():
    """sharding.DataParallel without a process group (world 1): hooked backward, bucket events, vsum_dp_finalize -- gradients
    and loss equal the plain step's (mse_with_mask_loss normalised by bs * Nmax, utils.py:55) up to the fp32 rounding of
    scaling after instead of before the backward."""
    from vsum_b200.model import SimNet
    from vsum_b200.sharding import DataParallel
    from vsum_b200.utils import mse_with_mask_loss
    lens, nmax = (140, 90, 33), 140
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.full((3, nmax, 1024), 1000.0, device="cuda")
    t = torch.full((3, nmax), 1000.0, device="cuda")
    for b, n in enumerate(lens):
        x[b, :n] = torch.rand((n, 1024), device="cuda", generator=g)
        t[b, :n] = torch.rand(n, device="cuda", generator=g)
    mask = x[:, :, 0] == 1000
    torch.manual_seed(21)
    model = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.0).cuda().train()
    model.train_precision = "fp32"      # scaling after instead of before the backward commutes up to fp32 rounding only in this mode
    pred, _ = model(x, mask)
    want_loss = mse_with_mask_loss(pred, t, mask)
    want_loss.backward()
    want = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.zero_grad(set_to_none=True)
    ddp = DataParallel(model, bucket_min_frames=0)
    pred, _ = model(x, mask)
    ddp.loss(pred, t, mask).backward()
    loss = ddp.finish()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(want_loss)) <= 1e-5 * abs(float(want_loss)) + 1e-9
    for k, p in model.named_parameters():
        torch.testing.assert_close(p.grad, want[k], rtol=2e-4, atol=1e-7 + 2e-5 * float(want[k].abs().max()))
    ddp.detach()
    assert model._dp is None


@pytest.mark.parametrize("fused", [True, False])
def test_optimizer_updates_reach_the_kernels(fused):
    """`torch.optim.Adam(fused=True)` updates the parameters without bumping their version counters: the differentiable
    forward must refresh the kernels' copy of the weights every step regardless (a stale copy trains on the initial weights
    for ever: constant loss), and an inference call after training must see the trained weights."""
    from vsum_b200.model import SimNet
    from vsum_b200.utils import mse_with_mask_loss
    torch.manual_seed(0)
    model = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.0).cuda().train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, fused=fused)
    T = 400
    g = torch.Generator(device="cuda").manual_seed(5)
    feats = torch.rand((T, 1024), device="cuda", generator=g)
    tgt = torch.rand((1, T), device="cuda", generator=g)
    cu = torch.tensor([0, 150, 400], dtype=torch.int32, device="cuda")
    nopad = torch.zeros((1, T), dtype=torch.bool, device="cuda")
    with torch.no_grad():
        before, _ = model.forward_packed(feats, cu, [150, 250])
        before = before.clone()
    losses = []
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        out, _ = model.forward_packed_train(feats, cu, [150, 250])
        loss = mse_with_mask_loss(out.view(1, T, 1), tgt, nopad)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert len(set(losses)) == 4, losses                      # every step saw the previous step's update
    with torch.no_grad():
        after, _ = model.forward_packed(feats, cu, [150, 250])
    assert not torch.equal(before, after)                      # inference after training runs on the trained weights
    want, _ = model.forward_packed_train(feats, cu, [150, 250])
    torch.testing.assert_close(after, want.detach(), rtol=2e-2, atol=2e-2)
