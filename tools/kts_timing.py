"""GPU box: kernel temporal segmentation of one video (n sub-sampled frames, up to m change points) on the GPU
against the CPU restatement (the reference's loops vectorised with numpy; the reference itself is slower)."""
import sys, time, numpy as np, torch
sys.path.insert(0, "video-summarization_b200"); sys.path.insert(0, ".")
from vsum_b200.segmentation import kts_segmentation
from oracle import kts_ref
for n, m in ((300, 30), (1000, 100), (2000, 200), (4000, 400)):
    rng = np.random.default_rng(n)
    x = rng.random((n, 256), dtype=np.float32)
    for c in range(0, n, n // 20): x[c:] += rng.random(256, dtype=np.float32)
    K = np.dot(x, x.T)
    Kd = torch.from_numpy(K).cuda()
    kts_segmentation(Kd, m, 1.0); torch.cuda.synchronize()
    t0 = time.time(); cps, costs = kts_segmentation(Kd, m, 1.0); torch.cuda.synchronize(); gpu = time.time() - t0
    line = f"n={n:5d} m={m:4d}: GPU {gpu * 1e3:8.1f} ms ({len(cps)} change points)"
    if n <= 1000:
        t0 = time.time(); want = kts_ref.kts_segmentation(K, m, 1.0); cpu = time.time() - t0
        line += f"   numpy restatement {cpu:7.1f} s   identical={np.array_equal(want[0], cps) and np.array_equal(want[1].view(np.int64), costs.view(np.int64))}"
    print(line, flush=True)
