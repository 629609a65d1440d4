"""GPU box: throughput of the data layer -- native collate into pinned memory + host-to-device copy -- on a pack
file of bench-shaped videos, against the padded path of the reference's collate (pad_sequence with the 1000 sentinel)."""
import os, sys, time, tempfile, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200.data import PackedDataset, PackedLoader, write_pack
from vsum_b200.synthetic import video_length

lens = [video_length(v, 128, 8192) for v in range(64)]
rng = np.random.default_rng(0)
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(d, "bench.vspack")
t0 = time.time()
write_pack(path, (dict(name=f"v{i}", features=rng.random((n, 1024), dtype=np.float32), gtscore=rng.random(n, dtype=np.float32)) for i, n in enumerate(lens)))
gb = sum(lens) * 4096 / 1e9
print(f"pack file: {len(lens)} videos, {sum(lens)} steps, {gb:.2f} GB written in {time.time() - t0:.1f} s")
ds = PackedDataset(path, split="train")
for threads in (1, 4, 16):
    loader = PackedLoader(ds, batch_size=16, collate_threads=threads)
    for _ in loader: pass                               # warm the page cache / pinned allocator
    torch.cuda.synchronize(); t0 = time.time()
    n = 0
    for b in loader: n += b.features.shape[0]
    torch.cuda.synchronize(); dt = time.time() - t0
    print(f"PackedLoader collate_threads={threads:2d}: {n * 4096 / dt / 1e9:6.2f} GB/s into HBM ({n / dt / 1e3:.0f} k steps/s)")
# reference-style: per-video tensors -> pad_sequence(padding_value=1000) -> .cuda()
feats = [torch.from_numpy(np.array(ds.array(i, 0))) for i in range(len(ds))]
torch.cuda.synchronize(); t0 = time.time()
n = 0
for s in range(0, len(feats), 16):
    x = torch.nn.utils.rnn.pad_sequence(feats[s:s + 16], batch_first=True, padding_value=1000).cuda()
    mask = x[:, :, 0] == 1000
    n += int((~mask).sum())
torch.cuda.synchronize(); dt = time.time() - t0
print(f"pad_sequence + .cuda() (reference collate): {n * 4096 / dt / 1e9:6.2f} GB/s of useful rows into HBM")
os.remove(path)
