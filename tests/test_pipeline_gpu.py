"""GPU: the whole path (features -> scorer -> sigmoid -> shot pooling -> knapsack -> mask -> F)
through the batched pipeline.  Given the scores the GPU produced, everything after the scorer is
checked bit-exactly against the oracle; the scores themselves are checked in test_scorer_*."""
import numpy as np
import pytest
import torch

from conftest import bits_equal
from oracle import c_oracle
from vsum_b200.model import SimNet
from vsum_b200.pipeline import DeviceBatch, Summarizer, pack_videos
from vsum_b200.synthetic import make_video, video_length

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_end_to_end_against_oracle(seeded_model_kwargs, precision):
    torch.manual_seed(1234)
    model = SimNet(**seeded_model_kwargs).cuda().eval()
    model.precision = precision
    vids = [make_video(800 + i, video_length(800 + i, 60, 1500), n_users=20) for i in range(24)]
    hb = pack_videos(vids)
    summ = Summarizer(model, "avg")
    out = summ.run_device(DeviceBatch(hb), want_intermediates=True)
    torch.cuda.synchronize()
    scores, f = out["scores"].cpu().numpy(), out["f"].cpu().numpy()
    assert np.all((scores > 0) & (scores < 1))
    for pos, idx in enumerate(hb.order):
        v = vids[idx]
        sc = scores[hb.cu_steps[pos]:hb.cu_steps[pos + 1]]
        want = c_oracle.video(sc, v.picks, v.n_frames, v.change_points, v.user_summary, "avg")
        assert bits_equal(np.float64(f[pos]), np.float64(want["f"])), (pos, idx)
        s0, s1 = hb.meta.cu_shots[pos], hb.meta.cu_shots[pos + 1]
        assert bits_equal(out["selected"][s0:s1].cpu().numpy(), want["selected"])
    f_host = summ.run_host(hb)                                # host buffers in, original order out
    assert bits_equal(f_host[hb.order], f)
    assert bits_equal(np.float64(np.mean(f_host)), np.float64(np.mean(f[np.argsort(hb.order)])))


@pytest.mark.parametrize("dataset,n_videos,n_users,method,lo,hi,test_size", [("summe", 25, 15, "max", 60, 650, 5),
                                                                              ("tvsum", 50, 20, "avg", 170, 1300, 10)])
def test_dsnet_split_shapes_against_oracle(seeded_model_kwargs, dataset, n_videos, n_users, method, lo, hi, test_size):
    """BASELINE config 2: SumMe-like (25 videos, 15 users, 'max') and TVSum-like (50 videos, 20 users, 'avg') sets, 5-fold
    test splits of the canonical / augmented sizes (src/splits_dsnet/*.yaml: 5 and 10 test keys), each split through
    `Summarizer.run_host`: given the scores the GPU produced, shot selections, keyshot masks and per-video F are the
    oracle's bit for bit, and so is the split mean (compute_metrics.py:92)."""
    torch.manual_seed(1234)
    model = SimNet(**seeded_model_kwargs).cuda().eval()
    base = 3000 if dataset == "summe" else 3100
    vids = [make_video(base + i, video_length(base + i, lo, hi), n_users=n_users) for i in range(n_videos)]
    summ = Summarizer(model, method)
    rng = np.random.default_rng(7)
    for fold in range(5):
        test = [vids[i] for i in rng.permutation(n_videos)[:test_size]]
        hb = pack_videos(test)
        out = summ.run_device(DeviceBatch(hb), want_intermediates=True)
        torch.cuda.synchronize()
        scores, f = out["scores"].cpu().numpy(), out["f"].cpu().numpy()
        masks, cu_frames = out["summary"].cpu().numpy(), hb.meta.sum_offsets
        want_f = np.empty(len(test))
        for pos, idx in enumerate(hb.order):
            v = test[idx]
            ref = c_oracle.video(scores[hb.cu_steps[pos]:hb.cu_steps[pos + 1]], v.picks, v.n_frames, v.change_points, v.user_summary, method)
            s0, s1 = hb.meta.cu_shots[pos], hb.meta.cu_shots[pos + 1]
            assert bits_equal(out["selected"][s0:s1].cpu().numpy(), ref["selected"]), (dataset, fold, pos)
            assert bits_equal(masks[cu_frames[pos]:cu_frames[pos + 1]].astype(np.int8), ref["summary"].astype(np.int8)), (dataset, fold, pos)
            assert bits_equal(np.float64(f[pos]), np.float64(ref["f"])), (dataset, fold, pos)
            want_f[idx] = ref["f"]
        f_host = summ.run_host(hb)                            # end-to-end call, caller's video order
        assert bits_equal(f_host, want_f)
        assert bits_equal(np.float64(np.mean(f_host)), np.float64(np.mean(want_f)))


def test_pipelined_submission_matches_single_stream(seeded_model_kwargs):
    torch.manual_seed(1234)
    model = SimNet(**seeded_model_kwargs).cuda().eval()
    vids = [make_video(850 + i, video_length(850 + i, 60, 1200), n_users=20) for i in range(12)]
    hb = pack_videos(vids)
    summ = Summarizer(model, "avg")
    want = summ.run_device(DeviceBatch(hb)).cpu().numpy()
    db = DeviceBatch(hb)
    outs = [summ.submit_device(db, i & 1) for i in range(4)]
    summ.drain()
    torch.cuda.synchronize()
    for f in outs:
        assert bits_equal(f.cpu().numpy(), want)
    dbe = [DeviceBatch(hb, pin_meta=True) for _ in range(2)]
    hosts = [summ.submit_host(dbe[i & 1], 2 + (i & 1)) for i in range(4)]
    summ.drain()
    torch.cuda.synchronize()
    assert bits_equal(hosts[-1].numpy(), want) and bits_equal(hosts[-2].numpy(), want)


def test_packed_loader_feeds_the_scorer(tmp_path):
    """Data layer end to end: pack file -> native collate into pinned staging -> async copy -> packed forward.
    Same scores as SimNet.forward on the padded batch the reference's collate_fn_train builds (dataset.py:157-161)."""
    import numpy as np
    import torch
    from vsum_b200.data import PackedDataset, PackedLoader, write_pack
    from vsum_b200.model import SimNet
    from vsum_b200.synthetic import make_video
    vids = [make_video(1500 + i, n, n_users=2) for i, n in enumerate((120, 64, 257, 90, 33))]
    path = str(tmp_path / "d.vspack")
    write_pack(path, [dict(name=v.name, features=v.features, gtscore=v.gtscore) for v in vids])
    torch.manual_seed(5)
    model = SimNet(num_heads=4, d_model=256, num_layers=2, sparsity=0., dropout=0.).cuda().eval()
    loader = PackedLoader(PackedDataset(path, split="train"), batch_size=3, collate_threads=4)
    seen = 0
    for b in loader:
        assert b.features.is_cuda and b.cu_seqlens.is_cuda and b.features.shape[0] == sum(b.seqlens)
        scores, _ = model.forward_packed(b.features, b.cu_seqlens, b.seqlens)
        nmax = max(b.seqlens)
        x = torch.full((len(b.ids), nmax, 1024), 1000.0)
        for k, i in enumerate(b.ids):
            x[k, :b.seqlens[k]] = torch.from_numpy(vids[i].features)
        x = x.cuda()
        mask = x[:, :, 0] == 1000
        with torch.no_grad():                                   # same (inference) kernels as forward_packed
            padded, _ = model(x, mask)
        want = torch.cat([padded[k, :n, 0] for k, n in enumerate(b.seqlens)])
        assert torch.equal(scores[:, 0], want)
        tgt = np.concatenate([vids[i].gtscore for i in b.ids])
        assert np.array_equal(b.targets.cpu().numpy(), tgt)
        seen += len(b.ids)
    assert seen == 5


def test_val_split_of_a_byte_pack_scores_like_the_float32_one(tmp_path):
    """`val` records of a pack written with user_summary_u8 keep the byte rows all the way into the overlap kernel;
    the F-scores are the float32 pack's, bit for bit (src/train.py:134-152 on either file)."""
    import numpy as np
    from conftest import bits_equal
    from vsum_b200.data import PackedDataset, write_pack
    from vsum_b200.evaluation import eval_fscores
    from vsum_b200.synthetic import make_scores, make_video
    vids = [make_video(2600 + i, n, n_users=5) for i, n in enumerate((150, 61, 333, 412))]
    recs = [dict(name=v.name, features=v.features, gtscore=v.gtscore, picks=v.picks, change_points=v.change_points,
                 n_frames=v.n_frames, user_summary=v.user_summary) for v in vids]
    got = []
    for u8 in (False, True):
        path = str(tmp_path / ("u8.vspack" if u8 else "f32.vspack"))
        write_pack(path, recs, user_summary_u8=u8)
        ds = PackedDataset(path, split="val")
        data, users = {}, {}
        for i, v in enumerate(vids):
            _, _, user = ds[i]
            assert user.user_summary.dtype == (np.uint8 if u8 else np.float32)
            data[user.name], users[user.name] = make_scores(2600 + i, v.n_steps), user
        got.append(np.asarray(eval_fscores(data, users)))
    assert len(got[0]) == 4 and bits_equal(got[0], got[1])


@pytest.mark.parametrize("u8,bf16", [(False, False), (True, True)])
def test_eval_loader_batches_match_host_packed_batches(tmp_path, seeded_model_kwargs, u8, bf16):
    """src/train.py:139-148 for many videos per step, fed from a page-locked pack: every `EvalBatch` of
    `PackedEvalLoader` (native metadata gather + one DMA per video out of the dataset) gives, bit for bit, the scores and
    F-scores of the same videos packed on the host (`pack_videos` + `DeviceBatch`), through the pipelined Summarizer and
    across the loader's slot ring (more batches than slots, two passes)."""
    from vsum_b200.data import PackedDataset, PackedEvalLoader, write_pack
    torch.manual_seed(1234)
    model = SimNet(**seeded_model_kwargs).cuda().eval()
    vids = [make_video(3100 + i, video_length(3100 + i, 40, 900), n_users=6) for i in range(22)]
    if bf16:                                                    # what a features_bf16 pack holds: the rounded rows
        for v in vids:
            v.features = torch.from_numpy(v.features).bfloat16().float().numpy()
    path = str(tmp_path / "val.vspack")
    write_pack(path, [dict(name=v.name, features=v.features, gtscore=v.gtscore, picks=v.picks, change_points=v.change_points,
                           n_frames=v.n_frames, user_summary=v.user_summary) for v in vids], user_summary_u8=u8, features_bf16=bf16)
    ds = PackedDataset(path, split="val", resident="pinned")
    loader = PackedEvalLoader(ds, batch_size=5, slots=2, cycle=True)
    assert len(loader) == 5
    summ = Summarizer(model, "avg")
    want = []
    for s in range(0, 22, 5):
        chunk = vids[s:s + 5]
        hb = pack_videos(chunk)
        if bf16:
            hb.features = hb.features.bfloat16()
        if u8:
            hb.meta.user_summary = hb.meta.user_summary.astype(np.uint8)
        out = summ.run_device(DeviceBatch(hb), want_intermediates=True)
        want.append((out["scores"].cpu().numpy(), out["f"].cpu().numpy(), [v.name for v in (chunk[i] for i in hb.order)]))
    torch.cuda.synchronize()
    got_f, names = [], []
    it = iter(loader)
    for k in range(10):                                         # two passes over the five batches
        b = next(it)
        names.append(b.names)
        assert b.features.dtype == (torch.bfloat16 if bf16 else torch.float32)
        got_f.append(summ.submit_device(b, k & 1, to_host=True))
        if k % 2 == 1:                                          # pinned landing buffers are per slot: read them before reuse
            summ.drain()
            torch.cuda.synchronize()
            got_f[-2], got_f[-1] = got_f[-2].numpy().copy(), got_f[-1].numpy().copy()
    it.close()
    for k in range(10):
        sc, f, nm = want[k % 5]
        assert names[k] == nm
        assert bits_equal(got_f[k], f), k
    assert len(loader.collate_ms) == 10 and ds.resident == "pinned"
