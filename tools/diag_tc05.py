"""Diagnostic (GPU box): runs the tcgen05 attention kernel on one 128-frame video with alternative
V-operand descriptors (VSUM_ATTN_V_DESC=lbo,sbo,kstep) in subprocesses and prints the error of each
against a PyTorch reference.  Used once while bringing the MN-major descriptor up."""
import os
import subprocess
import sys

CODE = r'''
import sys, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
lens = [128, 200]
T = sum(lens)
g = torch.Generator(device="cuda").manual_seed(0)
qkv = torch.randn((T, 768), device="cuda", generator=g).bfloat16()
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
scratch = torch.zeros(8 * (T // 128 + len(lens)) + 16, dtype=torch.int32, device="cuda")
_cabi.check(_cabi.load().vsum_debug_attention_tc05(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream), "attn")
torch.cuda.synchronize()
ref = torch.empty((T, 256), device="cuda"); off = 0
for n in lens:
    x = qkv[off:off+n].float()
    q, k, v = (x[:, i*256:(i+1)*256].view(n, 4, 64).permute(1, 0, 2) for i in range(3))
    ref[off:off+n] = (torch.softmax(q @ k.transpose(1, 2) / 16.0, -1) @ v).permute(1, 0, 2).reshape(n, 256); off += n
err = (out.float() - ref).abs()
print("max_err %.4f mean_err %.5f row0 %.4f nan %d" % (err.max().item(), err.mean().item(), err[0].max().item(), int(torch.isnan(out.float()).sum())))
'''
for desc in ("16,1024,2048", "1024,1024,2048", "2048,1024,2048", "16,2048,1024", "1024,2048,1024", "16,1024,256"):
    env = dict(os.environ, VSUM_ATTN_V_DESC=desc)
    try:
        r = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True, timeout=120)
        print(desc, "->", (r.stdout.strip() or r.stderr.strip()[-300:]))
    except subprocess.TimeoutExpired:
        print(desc, "-> timeout")
