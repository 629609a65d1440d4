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


@pytest.mark.parametrize("lens", [[128], [1], [37], [129], [256, 64], [300, 1, 127, 128, 513], [2048], [1000, 3000],
                                  [300] * 150, [8192, 4000, 77], [1] * 700, [129, 128, 127] * 60])
def test_attention(lens):   # the last four have more work items than resident CTAs (persistent loop)
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn((T, 768), device="cuda", generator=g).bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
    scratch = torch.zeros(2 * (T // 128 + len(lens)) + 1, dtype=torch.int32, device="cuda")
    _cabi.check(_cabi.load().vsum_debug_attention_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(),
                                                       scratch.data_ptr(), _stream()), "vsum_debug_attention_tc05")
    torch.cuda.synchronize()
    torch.testing.assert_close(out.float(), attention_ref(qkv, lens), rtol=2e-2, atol=2e-2)
