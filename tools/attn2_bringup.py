"""GPU box: correctness + timing of the attention forward kernels, for the shipped library and for every variant library
under tools/variants/ (built by tools/build_attn2_variants.sh with different -D switches).

  python tools/attn2_bringup.py [--quick] [--libs path ...]

Each library is loaded on its own (ctypes), checked against fp32 softmax attention on ragged cases, then timed on
148 x 2048 frames (one launch) and on the default bench workload's 256 log-uniform lengths.  Kernel version 1 (one tile
per CTA) is timed from the shipped library as the yardstick."""
import argparse, ctypes as C, glob, os, sys, time
import numpy as np, torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "video-summarization_b200"))
from vsum_b200.synthetic import video_length

vp = C.c_void_p
PRESCALED, PRE_SCALE = True, 1.0 / 1.4426950408889634     # the form the scorer runs: log2(e) / 16 folded into Q


def ref(qkv, lens):
    out = torch.empty((qkv.shape[0], 256), device="cuda"); off = 0
    for n in lens:
        x = qkv[off:off + n].float()
        q, k, v = (x[:, i * 256:(i + 1) * 256].view(n, 4, 64).permute(1, 0, 2) for i in range(3))
        out[off:off + n] = (torch.softmax(q @ k.transpose(1, 2) * (0.6931471805599453 if PRESCALED else 1.0 / 16.0), dim=-1) @ v).permute(1, 0, 2).reshape(n, 256)
        off += n
    return out


def setup(lens, entry, scale_k=None):
    T = sum(lens)
    g = torch.Generator(device="cuda").manual_seed(T)
    qkv = torch.randn((T, 768), device="cuda", generator=g)
    if PRESCALED:
        qkv[:, :256] *= 1.4426950408889634 / 16.0
    if scale_k is not None:
        qkv[:, 256:512] *= scale_k(T)
    qkv = qkv.bfloat16()
    cu = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device="cuda")
    out = torch.zeros((T, 256), dtype=torch.bfloat16, device="cuda")
    scratch = torch.zeros(8 * (T // 128 + len(lens)) + 16, dtype=torch.int32, device="cuda")

    def call():
        rc = entry(qkv.data_ptr(), cu.data_ptr(), len(lens), T, out.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    return qkv, out, call


def timed(call, reps=5):
    for _ in range(2): call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): call()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


CASES = [[128], [1], [37], [129], [256], [257], [256, 64], [300, 1, 127, 128, 513], [2048], [1000, 3000],
         [300] * 150, [8192, 4000, 77], [1] * 700, [129, 128, 127] * 60]
QUICK = [[128], [37], [129], [257], [300, 1, 127, 128, 513], [1000, 3000]]


def run_lib(path, version, quick, label=None):
    L = C.CDLL(path)
    entry0 = L.vsum_debug_attention_scaled_tc05
    entry0.argtypes = [vp, vp, C.c_int32, C.c_int64, C.c_float, vp, vp, vp]; entry0.restype = C.c_int
    entry = lambda a, b, c, d, *r: entry0(a, b, c, d, PRE_SCALE if PRESCALED else 1.0 / 16.0, *r)
    L.vsum_set_attention_kernel.argtypes = [C.c_int32]
    assert L.vsum_set_attention_kernel(version) == 0
    name = label or f"{os.path.basename(path)} v{version}"
    worst = 0.0
    for lens in (QUICK if quick else CASES):
        qkv, out, call = setup(lens, entry)
        call(); torch.cuda.synchronize()
        err = (out.float() - ref(qkv, lens)).abs().max().item()
        worst = max(worst, err)
        if not err < 2e-2:
            print(f"{name}: MISMATCH lens={lens[:6]}{'...' if len(lens) > 6 else ''} max err {err}", flush=True)
    # reference moves: later keys score much higher / everything scores far below zero
    for tag, fn in (("rising", lambda T: torch.linspace(0.2, 12.0, T, device="cuda")[:, None]),):
        lens = [640, 300]
        qkv, out, call = setup(lens, entry, fn)
        call(); torch.cuda.synchronize()
        err = (out.float() - ref(qkv, lens)).abs().max().item()
        worst = max(worst, err)
        if not err < 3e-2:
            print(f"{name}: MISMATCH ({tag} scores) max err {err}", flush=True)
    lens = [2048] * 148
    _, _, call = setup(lens, entry)
    ms = timed(call)
    tf = sum(4.0 * n * n * 256 for n in lens) / ms / 1e9
    lens_w = sorted((video_length(v, 128, 8192) for v in range(256)), reverse=True)
    _, _, call = setup(lens_w, entry)
    ms_w = timed(call)
    tf_w = sum(4.0 * n * n * 256 for n in lens_w) / ms_w / 1e9
    print(f"{name}: max err {worst:.4f} | 148x2048 {ms:.3f} ms {tf:.0f} TFLOP/s | bench workload {ms_w:.3f} ms {tf_w:.0f} TFLOP/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--libs", nargs="*")
    a = ap.parse_args()
    torch.zeros(1, device="cuda")
    shipped = os.path.join(os.path.dirname(__file__), "..", "video-summarization_b200", "vsum_b200", "libvsum_b200.so")
    libs = a.libs if a.libs else [shipped] + sorted(glob.glob(os.path.join(os.path.dirname(__file__), "variants", "*.so")))
    run_lib(shipped, 1, a.quick, "shipped v1 (one tile per CTA)")
    for p in libs:
        t0 = time.time()
        run_lib(p, 2, a.quick)
        if "timing" not in p:
            run_lib(p, 3, a.quick)
