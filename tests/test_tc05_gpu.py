"""GPU: the two tcgen05 kernels on their own (through the diagnostic C-ABI entry points) against
plain PyTorch fp32 references of the same op computed from the same bf16-rounded operands."""
import numpy as np
import pytest
import torch

from vsum_b200 import _cabi

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def gemm(A, W, bias, epi, residual=None, gamma=None, beta=None):
    M, K = A.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device="cuda")
    p = lambda t: None if t is None else t.data_ptr()
    _cabi.check(_cabi.load().vsum_debug_gemm_tc05(A.data_ptr(), W.data_ptr(), bias.data_ptr(), p(residual), p(gamma), p(beta),
                                                  out.data_ptr(), M, N, K, int(A.dtype == torch.float32), epi, _stream()),
                "vsum_debug_gemm_tc05")
    torch.cuda.synchronize()
    return out.float()


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (300, 768, 256), (1000, 1024, 256), (257, 256, 1024), (40000, 768, 256), (40000, 256, 256)])
@pytest.mark.parametrize("epi", [0, 1])
def test_gemm_bf16(M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    want = A.float() @ W.float().t() + bias
    if epi == 1:
        want = want.relu()
    got = gemm(A, W, bias, epi)
    torch.testing.assert_close(got, want.bfloat16().float(), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M", [64, 129, 5000])
def test_gemm_tf32_features(M):
    g = torch.Generator(device="cuda").manual_seed(M)
    A = torch.rand((M, 1024), device="cuda", generator=g)
    W = torch.randn((256, 1024), device="cuda", generator=g) / 32
    bias = torch.randn(256, device="cuda", generator=g)
    want = A.double() @ W.double().t() + bias.double()
    got = gemm(A, W, bias, 0)
    torch.testing.assert_close(got.double(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("M,K", [(128, 256), (333, 256), (40000, 256), (333, 1024)])
def test_gemm_residual_layernorm(M, K):
    g = torch.Generator(device="cuda").manual_seed(M + K)
    A = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    W = (torch.randn((256, K), device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(256, device="cuda", generator=g)
    res = torch.randn((M, 256), device="cuda", generator=g).bfloat16()
    gamma = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    beta = 0.1 * torch.randn(256, device="cuda", generator=g)
    want = torch.nn.functional.layer_norm(A.float() @ W.float().t() + bias + res.float(), (256,), gamma, beta)
    got = gemm(A, W, bias, 3, res, gamma, beta)
    torch.testing.assert_close(got, want, rtol=2e-2, atol=3e-2)


@pytest.mark.parametrize("M", [128, 1, 127, 333, 4096, 40000, 148 * 128 * 3 + 5])
def test_fused_ffn(M):
    """simnet.py:180-183 + 109-110 in one kernel (vsum_ffn_tc05.cu): LayerNorm(relu(x W1^T + b1) W2^T + b2 + x) against
    PyTorch fp32 on the same bf16-rounded operands, with the hidden activation rounded to bf16 where the kernel rounds
    it (it is the bf16 A operand of the second product).  Ragged M: partial last tile, fewer tiles than CTAs, several
    tiles per CTA."""
    g = torch.Generator(device="cuda").manual_seed(M)
    x = torch.randn((M, 256), device="cuda", generator=g).bfloat16()
    w1 = (torch.randn((1024, 256), device="cuda", generator=g) / 16).bfloat16()
    w2 = (torch.randn((256, 1024), device="cuda", generator=g) / 32).bfloat16()
    b1 = torch.randn(1024, device="cuda", generator=g) * 0.5
    b2 = torch.randn(256, device="cuda", generator=g)
    gamma = 1 + 0.1 * torch.randn(256, device="cuda", generator=g)
    beta = 0.1 * torch.randn(256, device="cuda", generator=g)
    hid = (x.float() @ w1.float().t() + b1).relu().bfloat16().float()
    want = torch.nn.functional.layer_norm(hid @ w2.float().t() + b2 + x.float(), (256,), gamma, beta)
    out = torch.full((M, 256), float("nan"), device="cuda").bfloat16()
    L = _cabi.load()
    _cabi.check(L.vsum_debug_ffn_tc05(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(),
                                      beta.data_ptr(), out.data_ptr(), M, torch.cuda.current_stream().cuda_stream), "vsum_debug_ffn_tc05")
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=3e-2)
    # and the two-launch form it replaces agrees to bf16 rounding of the output
    two = gemm(gemm(x, w1, b1, 1).bfloat16(), w2, b2, 3, x, gamma, beta)
    torch.testing.assert_close(out.float(), two, rtol=2e-2, atol=3e-2)


def attention_ref(qkv, lens):
    out = torch.empty((qkv.shape[0], 256), device="cuda")
    off = 0
    for n in lens:
        x = qkv[off:off + n].float()
        q, k, v = (x[:, i * 256:(i + 1) * 256].view(n, 4, 64).permute(1, 0, 2) for i in range(3))
        p = torch.softmax(q @ k.transpose(1, 2) / 16.0, dim=-1)
        out[off:off + n] = (p @ v).permute(1, 0, 2).reshape(n, 256)
        off += n
    return out


@pytest.fixture(params=[2, 1, 3], ids=["two_tile_persistent", "one_tile_per_cta", "two_tile_two_threads_per_row"])
def attn_kernel(request):
    """The forward kernels behind vsum_set_attention_kernel: 2 / 3 = persistent two-tile kernel with one / two softmax threads
    per query row (3 only differs on the pre-scaled inference fast pass), 1 = one tile per CTA."""
    L = _cabi.load()
    _cabi.check(L.vsum_set_attention_kernel(request.param), "vsum_set_attention_kernel")
    yield request.param
    _cabi.check(L.vsum_set_attention_kernel(2), "vsum_set_attention_kernel")


@pytest.mark.parametrize("lens", [[128], [1], [37], [129], [256], [257], [256, 64], [300, 1, 127, 128, 513], [2048], [1000, 3000],
                                  [300] * 150, [8192, 4000, 77], [1] * 700, [129, 128, 127] * 60])
def test_attention(lens, attn_kernel):   # the last four have more work items than resident CTAs (persistent loop)
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn((T, 768), device="cuda", generator=g).bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
    scratch = torch.zeros(_cabi.load().vsum_attention_scratch_ints(T, len(lens)), dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().vsum_debug_attention_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(),
                                                       scratch.data_ptr(), _stream()), "vsum_debug_attention_tc05")
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float(), attention_ref(qkv, lens), rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("lens", [[37], [1], [64], [65], [191], [300, 1, 127, 128, 513], [2048], [129, 128, 127] * 20, [8192, 4000, 77],
                                  [1] * 700, [300] * 150])
def test_attention_prescaled_q(lens, attn_kernel):
    """The form the scorer runs: d_model^-0.5 * log2(e) folded into Q (vsum_scorer_load_weights), scale = 1 / log2(e), so that
    the scores are base-2 exponents and the two-tile kernel exponentiates them as they are."""
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(T + 3)
    qkv = torch.randn((T, 768), device="cuda", generator=g)
    qkv[:, :256] *= 1.4426950408889634 / 16.0
    qkv = qkv.bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
    scratch = torch.zeros(_cabi.load().vsum_attention_scratch_ints(T, len(lens)), dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().vsum_debug_attention_scaled_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, 1.0 / 1.4426950408889634,
                                                              out.data_ptr(), scratch.data_ptr(), _stream()), "vsum_debug_attention_scaled_tc05")
    torch.cuda.synchronize()
    want, off = torch.empty((T, 256), device="cuda"), 0
    for n in lens:
        x = qkv[off:off + n].float()
        q, k, v = (x[:, i * 256:(i + 1) * 256].view(n, 4, 64).permute(1, 0, 2) for i in range(3))
        want[off:off + n] = (torch.softmax(q @ k.transpose(1, 2) * 0.6931471805599453, dim=-1) @ v).permute(1, 0, 2).reshape(n, 256)
        off += n
    torch.testing.assert_close(out.float(), want, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("case", ["rising", "falling", "below_zero", "far_below_zero", "far_above_zero", "one_hot"])
def test_attention_exponent_reference_moves(case, attn_kernel):
    """Score distributions that push the running exponent reference of the softmax around: the two-tile kernel
    exponentiates against a reference that only moves when a tile maximum leaves a +-24 (log2) window and then
    recomputes the tile; the result must not depend on any of that."""
    lens = [640, 300, 129]
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(11)
    qkv = torch.randn((T, 768), device="cuda", generator=g)
    if case == "rising":          # later keys score much higher: the reference climbs tile after tile
        qkv[:, 256:512] *= torch.linspace(0.2, 14.0, T, device="cuda")[:, None]
    elif case == "falling":       # the first tile holds the maximum
        qkv[:, 256:512] *= torch.linspace(14.0, 0.2, T, device="cuda")[:, None]
    elif case == "below_zero":       # every logit around -60 log2 units: the reference moves down after the first tile
        qkv[:, 0:256] = 0.3 * qkv[:, 0:256] + 3.3
        qkv[:, 256:512] = 0.3 * qkv[:, 256:512] - 3.3
    elif case == "far_below_zero":   # every logit around -140 log2 units: exp2(s) alone underflows (exact pass)
        qkv[:, 0:256] = 0.3 * qkv[:, 0:256] + 5.0
        qkv[:, 256:512] = 0.3 * qkv[:, 256:512] - 5.0
    elif case == "far_above_zero":   # every logit around +90 log2 units
        qkv[:, 0:256] = 0.3 * qkv[:, 0:256] + 5.0
        qkv[:, 256:512] = 0.3 * qkv[:, 256:512] + 5.0
    else:                         # one key per row dominates by a huge margin
        qkv[:, 0:512] *= 6.0
    qkv = qkv.bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
    scratch = torch.zeros(_cabi.load().vsum_attention_scratch_ints(T, len(lens)), dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().vsum_debug_attention_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(),
                                                       scratch.data_ptr(), _stream()), "vsum_debug_attention_tc05")
    torch.cuda.synchronize()
    want = attention_ref(qkv, lens)
    assert torch.isfinite(out.float()).all()
    torch.testing.assert_close(out.float(), want, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("M,N,K,epi", [(300, 256, 1024, 5), (5000, 768, 256, 5), (777, 1024, 256, 6), (40000, 256, 256, 5)])
def test_gemm_tf32_fp32_output(M, N, K, epi):
    """Training-path linear: fp32 in memory, tf32 MMA, fp32 output."""
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn((M, K), device="cuda", generator=g)
    W = torch.randn((N, K), device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty((M, N), dtype=torch.float32, device="cuda")
    _cabi.check(_cabi.load().vsum_debug_gemm_tc05(A.data_ptr(), W.data_ptr(), bias.data_ptr(), None, None, None, out.data_ptr(),
                                                  M, N, K, 1, epi, _stream()), "vsum_debug_gemm_tc05")
    torch.cuda.synchronize()
    want = A.double() @ W.double().t() + bias.double()
    if epi == 6:
        want = want.relu()
    torch.testing.assert_close(out.double(), want, rtol=3e-3, atol=3e-3)


@pytest.mark.parametrize("bf16", [True])
@pytest.mark.parametrize("M,N,K", [(32, 128, 256), (300, 256, 256), (5000, 768, 256), (4097, 256, 1024), (70000, 1024, 256)])
def test_wgrad_tc05(M, N, K, bf16):
    """dW = dY^T X with both operands MN-major (contraction over frames), split over frames + atomics."""
    g = torch.Generator(device="cuda").manual_seed(M + K)
    dY = torch.randn((M, N), device="cuda", generator=g)
    X = torch.randn((M, K), device="cuda", generator=g)
    dW = torch.zeros((N, K), device="cuda")
    db = torch.zeros(N, device="cuda")
    scratch = torch.empty(M * (N + K), dtype=torch.bfloat16, device="cuda") if bf16 else None
    _cabi.check(_cabi.load().vsum_debug_wgrad_tc05(dY.data_ptr(), X.data_ptr(), dW.data_ptr(), db.data_ptr(), M, N, K,
                                                   None if scratch is None else scratch.data_ptr(), _stream()),
                "vsum_debug_wgrad_tc05")
    torch.cuda.synchronize()
    want = dY.double().t() @ X.double()
    scale = want.abs().max().item()
    err = (dW.double() - want).abs().max().item()
    assert err <= (1.5e-2 if bf16 else 3e-3) * scale, f"max err {err:.3e} of scale {scale:.3e}; dW[0,:4]={dW[0,:4].tolist()} want {want[0,:4].tolist()}"
    torch.testing.assert_close(db.double(), dY.double().sum(0), rtol=1e-4, atol=1e-3 * M ** 0.5)


def _drop_keep(seed, q_rows, h, keys, thresh):
    """numpy restatement of the grouped dropout draw (vsum_kernels.cuh: dropout_bits64 = 4 Philox-2x32 rounds,
    attn_drop_group_index)."""
    M32 = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        idx = ((q_rows.astype(np.uint64)[:, None] * np.uint64(4) + np.uint64(h)) << np.uint64(20)) \
            ^ (keys.astype(np.uint64)[None, :] >> np.uint64(2)) ^ np.uint64(0x5A5A000000000000)
        x0, x1 = idx & M32, idx >> np.uint64(32)
        k = (np.uint64(seed) & M32) ^ (((np.uint64(seed) >> np.uint64(32)) * np.uint64(0x85EBCA6B)) & M32)
        for _ in range(4):
            m = x0 * np.uint64(0xD256D193)
            x0, x1 = ((m >> np.uint64(32)) ^ k ^ x1) & M32, m & M32
            k = (k + np.uint64(0x9E3779B9)) & M32
        z = (x0 << np.uint64(32)) | x1
        lane = (keys.astype(np.uint64)[None, :] & np.uint64(3)) * np.uint64(16)
        return ((z >> lane) & np.uint64(0xFFFF)) >= np.uint64(thresh)


@pytest.mark.parametrize("lens,p", [([128], 0.0), ([37], 0.0), ([129, 300], 0.0), ([300, 1, 127, 513], 0.3), ([1000, 2100], 0.0),
                                    ([700, 260], 0.1)])
def test_attention_train_and_backward(lens, p, attn_kernel):
    """tcgen05 attention with log-sum-exp + dropout, and its tcgen05 backward, against torch autograd (fp32
    math on the same bf16-rounded operands, the same dropout mask)."""
    T, B, seed = sum(lens), len(lens), 0x1234567 + sum(lens)
    g = torch.Generator(device="cuda").manual_seed(T + 5)
    qkv = torch.randn((T, 768), device="cuda", generator=g).bfloat16()
    d_out = torch.randn((T, 256), device="cuda", generator=g).bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), device="cuda")                      # the training variant writes fp32
    lse2 = torch.zeros((T, 4), device="cuda")
    scratch = torch.zeros(_cabi.load().vsum_attention_scratch_ints(T, B), dtype=torch.int32, device="cuda")
    L = _cabi.load()
    _cabi.check(L.vsum_debug_attention_train_tc05(qkv.data_ptr(), cu.data_ptr(), B, T, out.data_ptr(), lse2.data_ptr(), p, seed,
                                                  scratch.data_ptr(), _stream()), "vsum_debug_attention_train_tc05")
    torch.cuda.synchronize()
    thresh = 0 if p <= 0 else int(p * 65536 + 0.5)
    ks = 65536.0 / (65536 - thresh)
    x = qkv.float().requires_grad_(True)
    want_out, want_lse, off = [], [], 0
    for n in lens:
        heads, lses = [], []
        for h in range(4):
            q, k, v = (x[off:off + n, i * 256 + h * 64: i * 256 + (h + 1) * 64] for i in range(3))
            s = (q @ k.t()) / 16.0
            lses.append(torch.logsumexp(s, dim=1) * 1.4426950408889634)
            pr = torch.softmax(s, dim=1)
            if thresh:
                keep = _drop_keep(seed, np.arange(off, off + n), h, np.arange(n), thresh)
                pr = pr * torch.from_numpy(keep).to(pr) * ks
            heads.append(pr @ v)
        want_out.append(torch.cat(heads, dim=1))
        want_lse.append(torch.stack(lses, dim=1))
        off += n
    want_out, want_lse = torch.cat(want_out), torch.cat(want_lse)
    torch.testing.assert_close(lse2, want_lse.detach(), rtol=1e-3, atol=2e-3)
    torch.testing.assert_close(out, want_out.detach(), rtol=2e-2, atol=2e-2)

    want_out.backward(d_out.float())
    delta = (out * d_out.float()).view(T, 4, 64).sum(-1).contiguous()
    dqkv = torch.full((T, 768), float("nan"), device="cuda")
    _cabi.check(L.vsum_debug_attention_bwd_tc05(qkv.data_ptr(), d_out.data_ptr(), lse2.data_ptr(), delta.data_ptr(), cu.data_ptr(),
                                                B, T, p, seed, dqkv.data_ptr(), scratch.data_ptr(), _stream()),
                "vsum_debug_attention_bwd_tc05")
    torch.cuda.synchronize()
    for i, name in enumerate(("dq", "dk", "dv")):
        got, want = dqkv[:, i * 256:(i + 1) * 256], x.grad[:, i * 256:(i + 1) * 256]
        scale = want.abs().max().item()
        err = (got - want).abs().max().item()
        assert err <= 2e-2 * scale, f"{name}: max err {err:.3e} vs scale {scale:.3e}"
