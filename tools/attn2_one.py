"""GPU box: the attention forward kernel alone on 148 x 2048 frames (or the bench workload's lengths with --workload),
a few launches -- the command profiled by ncu and the one that prints the -DVSUM_A2_TIMING phase stamps.
  python tools/attn2_one.py [--lib path] [--version 2] [--workload] [--reps 3]"""
import argparse, ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "video-summarization_b200"))
from vsum_b200.synthetic import video_length

ap = argparse.ArgumentParser()
ap.add_argument("--lib", default=os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "video-summarization_b200", "vsum_b200", "libvsum_b200.so"))
ap.add_argument("--version", type=int, default=2)
ap.add_argument("--workload", action="store_true")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--lens", default=None, help="python expression for the list of video lengths, e.g. \"[300]*150\"")
ap.add_argument("--check", action="store_true", help="compare with fp32 softmax attention")
ap.add_argument("--prescaled", action="store_true", help="fold log2(e)/16 into Q and run with scale 1/log2(e), as the scorer does")
a = ap.parse_args()
PRESCALED, PRE_SCALE = a.prescaled, 1.0 / 1.4426950408889634
vp = C.c_void_p
L = C.CDLL(a.lib)
entry0 = L.vsum_debug_attention_scaled_tc05
entry0.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_float, vp, vp, vp]; entry0.restype = C.c_int
entry = lambda a, b, c, d, *r: entry0(a, b, c, d, PRE_SCALE if PRESCALED else 1.0 / 16.0, *r)
L.vsum_set_attention_kernel.argtypes = [C.c_int32]
assert L.vsum_set_attention_kernel(a.version) == 0
lens = eval(a.lens) if a.lens else (sorted((video_length(v, 128, 8192) for v in range(256)), reverse=True) if a.workload else [2048] * 148)
T = sum(lens)
g = torch.Generator(device="cuda").manual_seed(T)
qkv = torch.randn((T, 768), device="cuda", generator=g)
if PRESCALED:
    qkv[:, :256] *= 1.4426950408889634 / 16.0
qkv = qkv.bfloat16()
cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
scratch = torch.zeros(8 * (T // 128 + len(lens)) + 16, dtype=torch.int32, device="cuda")
def call():
    rc = entry(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0, rc
call(); torch.cuda.synchronize()
if a.check:
    ref = torch.empty((T, 256), device="cuda"); off = 0
    for n in lens:
        x = qkv[off:off + n].float()
        q, k, v = (x[:, i * 256:(i + 1) * 256].view(n, 4, 64).permute(1, 0, 2) for i in range(3))
        ref[off:off + n] = (torch.softmax(q @ k.transpose(1, 2) * (0.6931471805599453 if PRESCALED else 1.0 / 16.0), dim=-1) @ v).permute(1, 0, 2).reshape(n, 256)
        off += n
    print("max err", (out.float() - ref).abs().max().item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps): call()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
print(f"{os.path.basename(a.lib)} v{a.version} {'workload' if a.workload else '148x2048'}: {ms:.3f} ms {sum(4.0 * n * n * 256 for n in lens) / ms / 1e9:.0f} TFLOP/s")
