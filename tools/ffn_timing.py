"""GPU box: fused feed-forward kernel against the two GEMM launches it replaces, on the bench batch's row count."""
import sys, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
L = _cabi.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 539635
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.randn((M, 256), device="cuda", generator=g).bfloat16()
w1 = (torch.randn((1024, 256), device="cuda", generator=g) / 16).bfloat16()
w2 = (torch.randn((256, 1024), device="cuda", generator=g) / 32).bfloat16()
b1 = torch.randn(1024, device="cuda", generator=g); b2 = torch.randn(256, device="cuda", generator=g)
gamma = torch.ones(256, device="cuda"); beta = torch.zeros(256, device="cuda")
out = torch.empty((M, 256), device="cuda", dtype=torch.bfloat16); hid = torch.empty((M, 1024), device="cuda", dtype=torch.bfloat16)
s = torch.cuda.current_stream().cuda_stream
def fused(): _cabi.check(L.vsum_debug_ffn_tc05(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), M, s), "ffn")
def two():
    _cabi.check(L.vsum_debug_gemm_tc05(x.data_ptr(), w1.data_ptr(), b1.data_ptr(), None, None, None, hid.data_ptr(), M, 1024, 256, 0, 1, s), "fc1")
    _cabi.check(L.vsum_debug_gemm_tc05(hid.data_ptr(), w2.data_ptr(), b2.data_ptr(), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), M, 256, 1024, 0, 3, s), "fc2")
flops = 2.0 * M * 256 * 1024 * 2
for name, fn in (("fused", fused), ("two launches", two), ("fused", fused)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name:14s} M={M}: {ms:.4f} ms  {flops / ms / 1e9:.0f} TFLOP/s")
