"""GPU parity: fp32 mode of the CUDA scorer against the reference's fp32 outputs (golden fixtures)
and the torch fp32 oracle.  Tolerance: 1e-5 (BASELINE.json, 'fp32 reference mode')."""
import numpy as np
import pytest
import torch

from oracle import scorer_ref
from vsum_b200.model import SimNet
from vsum_b200.synthetic import make_video

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-5, 1e-5


def seeded(kwargs, precision):
    torch.manual_seed(1234)
    m = SimNet(**kwargs).cuda().eval()
    m.precision = precision
    return m


@pytest.fixture(scope="module")
def model(seeded_model_kwargs):
    return seeded(seeded_model_kwargs, "fp32")


def test_golden_logits_and_feats(model, scorer_golden):
    for vid, n in [tuple(int(x) for x in r) for r in scorer_golden["cases"]]:
        x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0).cuda()
        with torch.no_grad():
            logits, feats = model(x)
        assert logits.shape == (1, n, 1) and feats.shape == (1, n, 256)
        np.testing.assert_allclose(logits.view(-1).cpu().numpy(), scorer_golden[f"logits_{vid}"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(feats[0, :4].cpu().numpy(), scorer_golden[f"feats_head_{vid}"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(feats[0, -4:].cpu().numpy(), scorer_golden[f"feats_tail_{vid}"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(feats[0].double().sum(1).cpu().numpy(), scorer_golden[f"feats_rowsum_{vid}"], rtol=1e-4, atol=2e-4)


def test_longer_than_reference_table(model, scorer_golden):
    x = torch.from_numpy(make_video(106, 2300).features).unsqueeze(0).cuda()
    with torch.no_grad():
        logits, _ = model(x)
    np.testing.assert_allclose(logits.view(-1).cpu().numpy(), scorer_golden["logits_106"], rtol=RTOL, atol=ATOL)


def test_full_length_videos_against_reference(model, scorer_long_golden):
    """N = 4096 and 8192 (BASELINE config 5's upper end) against the reference's own outputs (its PositionalEncoding class
    with maxlen=8192, simnet.py:220-238)."""
    for vid, n in [tuple(int(x) for x in r) for r in scorer_long_golden["cases"]]:
        x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0).cuda()
        with torch.no_grad():
            logits, feats = model(x)
        np.testing.assert_allclose(logits.view(-1).cpu().numpy(), scorer_long_golden[f"logits_{vid}"], rtol=RTOL, atol=ATOL)
        np.testing.assert_allclose(feats[0].double().sum(1).cpu().numpy(), scorer_long_golden[f"feats_rowsum_{vid}"], rtol=1e-4, atol=3e-4)


def test_masked_mse_against_reference_loss(scorer_golden):
    """`mse_with_mask_loss` (utils.py:45-56) through vsum_masked_mse on the reference's own padded-batch logits: the
    reference's loss value (tests/golden/make_golden.py: padded_loss) must come back, and the gradient must be the
    analytic one with the reference's bs * Nmax normalisation."""
    from vsum_b200.utils import mse_with_mask_loss
    lens = (300, 180, 77)
    logits = torch.full((3, 300, 1), 7.0)                                 # padded rows: arbitrary finite values, masked out
    tgt = torch.full((3, 300), 1000.0)
    mask = torch.ones((3, 300), dtype=torch.bool)
    for b, n in enumerate(lens):
        logits[b, :n, 0] = torch.from_numpy(scorer_golden[f"padded_logits_{b}"])
        tgt[b, :n] = torch.from_numpy(make_video(110 + b, n).gtscore)
        mask[b, :n] = False
    out = logits.cuda().requires_grad_(True)
    loss = mse_with_mask_loss(out, tgt.cuda(), mask.cuda())
    np.testing.assert_allclose(loss.item(), float(scorer_golden["padded_loss"]), rtol=2e-6, atol=0)
    loss.backward()
    want_grad = 2.0 * (logits[:, :, 0] - tgt) * (~mask) / (3 * 300)
    np.testing.assert_allclose(out.grad[:, :, 0].cpu().numpy(), want_grad.numpy(), rtol=1e-6, atol=1e-9)
    with pytest.raises(Exception):                                        # no CPU fallback
        mse_with_mask_loss(logits, tgt, mask)


def test_padded_batch_with_key_mask(model, scorer_golden):
    lens = (300, 180, 77)
    x = torch.full((3, 300, 1024), 1000.0)
    for b, n in enumerate(lens):
        x[b, :n] = torch.from_numpy(make_video(110 + b, n).features)
    x = x.cuda()
    mask = x[:, :, 0] == 1000                                            # train.py:118
    with torch.no_grad():
        logits, feats = model(x, mask)
    assert logits.shape == (3, 300, 1)
    for b, n in enumerate(lens):
        np.testing.assert_allclose(logits[b, :n, 0].cpu().numpy(), scorer_golden[f"padded_logits_{b}"], rtol=RTOL, atol=ATOL)
    # a non-tensor mask is ignored (simnet.py:38, train.py:162)
    with torch.no_grad():
        a, _ = model(x[:1, :77], True)
        b, _ = model(x[:1, :77])
    assert torch.equal(a, b)


def test_packed_equals_per_video(model):
    vids = [make_video(400 + i, n) for i, n in enumerate([64, 129, 1, 500])]
    feats = torch.from_numpy(np.concatenate([v.features for v in vids])).cuda()
    lens = [v.n_steps for v in vids]
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32).cuda()
    packed, _ = model.forward_packed(feats, cu, lens)
    off = 0
    for v in vids:
        with torch.no_grad():
            single, _ = model(torch.from_numpy(v.features).unsqueeze(0).cuda())
        np.testing.assert_allclose(packed[off:off + v.n_steps, 0].cpu().numpy(), single.view(-1).cpu().numpy(), rtol=1e-6, atol=1e-6)
        off += v.n_steps


def test_small_generic_model(scorer_golden):
    m = seeded(dict(num_heads=4, d_model=64, num_layers=2, sparsity=0., use_cls=False, dropout=0.1,
                    num_classes=1, use_pos=True), "fp32")
    with torch.no_grad():
        logits, feats = m(torch.from_numpy(make_video(120, 50).features).unsqueeze(0).cuda())
    np.testing.assert_allclose(logits.view(-1).cpu().numpy(), scorer_golden["small_logits"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(feats[0].cpu().numpy(), scorer_golden["small_feats"], rtol=RTOL, atol=ATOL)


def test_against_torch_oracle_random_weights():
    torch.manual_seed(7)
    m = SimNet(num_heads=4, d_model=256, num_layers=2, dropout=0.0).cuda().eval()
    m.precision = "fp32"
    with torch.no_grad():
        for p in m.parameters():           # non-trivial LayerNorm affine and biases
            p.add_(0.05 * torch.randn_like(p))
    sd = {k: v.cpu() for k, v in m.state_dict().items()}
    x = torch.from_numpy(make_video(500, 257).features).unsqueeze(0)
    want_logits, want_feats = scorer_ref.scorer_forward(sd, x, num_heads=4)
    with torch.no_grad():
        logits, feats = m(x.cuda())
    np.testing.assert_allclose(logits.cpu().numpy(), want_logits.numpy(), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(feats.cpu().numpy(), want_feats.numpy(), rtol=RTOL, atol=2e-5)


def test_error_paths(model):
    with pytest.raises(Exception):
        model(torch.zeros(1, 4, 1024))                                   # CPU tensor: no fallback
    with torch.no_grad():
        model.precision = "int8"
        with pytest.raises(ValueError):
            model(torch.zeros(1, 4, 1024, device="cuda"))
        model.precision = "fp32"
