"""GPU box: per-phase cycle breakdown of the attention softmax loop (debug .so built with
-DVSUM_ATTN_TIMING) and kernel time at N=2048, with 2 CTAs/SM and forced 1 CTA/SM."""
import os, sys, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
lens = [2048] * 148
T = sum(lens)
qkv = (torch.randn((T, 768), device="cuda") * 1.0).bfloat16()
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
scratch = torch.zeros(8 * (T // 128 + len(lens)) + 16, dtype=torch.int32, device="cuda")
L = _cabi.load()
def run():
    _cabi.check(L.vsum_debug_attention_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "attn")
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
fl = sum(4.0 * n * n * 256 for n in lens)
print(f"lib={os.path.basename(_cabi.LIB_PATH)} one_cta={os.environ.get('VSUM_ATTN_ONE_CTA','0')} attention N=2048 x148: {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s")
