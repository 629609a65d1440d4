"""GPU box: time vsum_rank_correlation (Kendall tau-b + Spearman rho for every user of every video) on
device-resident inputs, for the bench workload shape (N log-uniform, 20 users) and a TVSum-like one, and scipy
(the reference's implementation) on a few videos of the same shape on the host."""
import sys, time, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
from vsum_b200.synthetic import video_length

def run(lens, users, levels):
    B = len(lens)
    nfr = np.asarray([15 * (n - 1) + 8 for n in lens], dtype=np.int32)
    cu_steps = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    picks = np.concatenate([np.arange(0, f, 15, dtype=np.int32)[:n] for n, f in zip(lens, nfr)])
    us_off = np.concatenate([[0], np.cumsum(nfr.astype(np.int64) * users)]).astype(np.int64)
    cu_users = (np.arange(B + 1) * users).astype(np.int32)
    T, total_users, total = int(cu_steps[-1]), B * users, int(us_off[-1])
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(3)
    scores = torch.rand(T, device=dev, generator=g)
    us = torch.rand(total, device=dev, generator=g)
    if levels:
        us = torch.floor(us * 5) + 1
    t = lambda a: torch.from_numpy(a).to(dev)
    d = [t(cu_steps), t(picks), t(nfr), t(us_off), t(cu_users), t(nfr.copy())]
    L = _cabi.load()
    need = L.vsum_rank_correlation_workspace_bytes(total, T, B, total_users)
    ws = torch.empty(need + 1024, dtype=torch.uint8, device=dev)
    wp = (ws.data_ptr() + 1023) // 1024 * 1024
    out = torch.empty(2 * B, dtype=torch.float64, device=dev)
    def call():
        _cabi.check(L.vsum_rank_correlation(scores.data_ptr(), d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), us.data_ptr(), d[3].data_ptr(),
                                            d[4].data_ptr(), d[5].data_ptr(), B, T, max(lens), total_users, total, wp, need, out.data_ptr(),
                                            out[B:].data_ptr(), None, None, torch.cuda.current_stream().cuda_stream), "vsum_rank_correlation")
    call(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): call()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"videos={B} users={users} levels={levels} frames={int(nfr.sum())} user-score elements={total/1e6:.1f} M workspace={need/1e9:.2f} GB: "
          f"{ms:.2f} ms  ({total/ms/1e6:.2f} G elements/s, {B/ms*1e3:.0f} videos/s) mean tau {out[:B].nanmean().item():+.5f}", flush=True)
    return scores, us, cu_steps, picks, nfr, us_off

lens = [video_length(v, 128, 8192) for v in range(256)]
run(lens, 20, False)
run(lens, 20, True)
tv = [int(n) for n in np.random.default_rng(0).integers(200, 401, 50)]
scores, us, cu, picks, nfr, off = run(tv, 20, True)
# scipy on the host for the first 3 TVSum-like videos
from scipy import stats
sc, usn = scores.cpu().numpy(), us.cpu().numpy()
t0 = time.time()
for v in range(3):
    n, f = tv[v], int(nfr[v])
    fs = np.repeat(sc[cu[v]:cu[v + 1]], np.diff(np.concatenate([picks[cu[v]:cu[v + 1]], [f]])))
    pr = stats.rankdata(-fs)
    for u in range(20):
        row = usn[off[v] + u * f: off[v] + (u + 1) * f]
        ur = stats.rankdata(-row); stats.spearmanr(pr, ur); stats.kendalltau(pr, ur)
dt = time.time() - t0
print(f"scipy (1 core): {dt / 3 * 1e3:.1f} ms per TVSum-like video (20 users) -> {3 / dt:.1f} videos/s")
