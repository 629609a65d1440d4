"""CPU: the packed dataset file, the native memory-mapped reader and the padding-free collate
(vsum_pack_*, vsum_b200/data) against the arrays they were written from -- the records mirror
src/data/dataset.py:64-168."""
import numpy as np
import pytest
import torch

from vsum_b200 import _cabi
from vsum_b200.data import PackedDataset, PackedLoader, write_pack
from vsum_b200.synthetic import make_video


def _videos(ns, first=700):
    out = []
    for i, n in enumerate(ns):
        v = make_video(first + i, n, n_users=3, with_features=True, with_user_scores=True)
        out.append(dict(name=v.name, features=v.features, gtscore=v.gtscore, picks=v.picks, change_points=v.change_points,
                        n_frames=v.n_frames, user_summary=v.user_summary, user_scores=v.user_scores,
                        video_rep=np.random.default_rng(i).random(512, dtype=np.float32)))
    return out


@pytest.mark.parametrize("u8", [False, True])
def test_roundtrip_and_records(tmp_path, u8):
    vids = _videos((60, 7, 131, 300))
    path = str(tmp_path / "d.vspack")
    write_pack(path, vids, user_summary_u8=u8)
    ds = PackedDataset(path, split="val")
    assert len(ds) == 4 and ds.feature_dim == 1024 and ds.names == [v["name"] for v in vids]
    for i, v in enumerate(vids):
        feats, tgt, user = ds[i]
        assert torch.equal(feats, torch.from_numpy(v["features"])) and torch.equal(tgt, torch.from_numpy(v["gtscore"]))
        assert np.array_equal(user.picks, v["picks"]) and np.array_equal(user.change_points, v["change_points"])
        assert int(user.n_frames) == v["n_frames"] and user.name == v["name"]
        assert np.array_equal(np.asarray(user.user_summary, np.float32), v["user_summary"].astype(np.float32))
        assert user.user_summary.dtype == (np.uint8 if u8 else np.float32)
        assert np.array_equal(user.user_scores, v["user_scores"])
    tr = PackedDataset(path, split="train", min_steps=50)            # dataset.py:121 keeps videos with > 50 steps
    assert tr.names == [vids[0]["name"], vids[2]["name"], vids[3]["name"]] and len(tr[0]) == 2
    sel = PackedDataset(path, split="val", keys=["some/dir/" + vids[1]["name"], vids[3]["name"]])     # split keys, dataset.py:137-140
    assert sel.names == [vids[1]["name"], vids[3]["name"]]
    pre = PackedDataset(path, split="pretrain")
    f, rep = pre[2]
    assert rep.shape == (512,) and np.array_equal(rep.numpy(), vids[2]["video_rep"])


def test_bf16_feature_pack(tmp_path):
    """features_bf16: rows are the round-to-nearest-even bfloat16 of the float32 features (what torch's .bfloat16() gives),
    served to the val split only, and the collate gathers them as bf16 rows."""
    vids = _videos((60, 7, 131), first=810)
    vids[0]["features"] = vids[0]["features"].copy()
    vids[0]["features"][0, :4] = np.array([1.00390625, 1.01171875, -3.0e-39, 65504.0], np.float32)   # ties, a denormal, a large value
    path = str(tmp_path / "d.vspack")
    write_pack(path, vids, features_bf16=True, user_summary_u8=True)
    ds = PackedDataset(path, split="val")
    assert ds.features_bf16
    for i, v in enumerate(vids):
        feats, tgt, user = ds[i]
        assert feats.dtype == torch.bfloat16 and torch.equal(feats, torch.from_numpy(v["features"]).bfloat16())
        assert torch.equal(tgt, torch.from_numpy(v["gtscore"])) and user.user_summary.dtype == np.uint8
    b = next(iter(PackedLoader(ds, batch_size=3, device="cpu", collate_threads=3)))
    assert b.features.dtype == torch.bfloat16 and b.cu_seqlens.tolist() == [0, 60, 67, 198]
    assert torch.equal(b.features, torch.cat([torch.from_numpy(v["features"]).bfloat16() for v in vids]))
    with pytest.raises(ValueError, match="inference"):
        PackedDataset(path, split="train")


def test_byte_user_summaries_must_be_lossless(tmp_path):
    vids = _videos((40,))
    vids[0]["user_summary"] = vids[0]["user_summary"] * 0.5          # not representable as bytes
    with pytest.raises(ValueError, match="lossless"):
        write_pack(str(tmp_path / "d.vspack"), vids, user_summary_u8=True)


@pytest.mark.parametrize("threads", [1, 5])
def test_collate_is_packed_and_ordered(tmp_path, threads):
    vids = _videos((60, 7, 131, 300, 1, 2000), first=720)
    path = str(tmp_path / "d.vspack")
    write_pack(path, vids)
    ds = PackedDataset(path, split="train")
    loader = PackedLoader(ds, batch_size=4, device="cpu", collate_threads=threads)
    batches = list(loader)
    assert len(loader) == 2 and [b.ids for b in batches] == [[0, 1, 2, 3], [4, 5]]
    for b in batches:
        want = np.concatenate([vids[i]["features"] for i in b.ids])
        assert np.array_equal(b.features.numpy(), want)
        assert np.array_equal(b.targets.numpy(), np.concatenate([vids[i]["gtscore"] for i in b.ids]))
        assert b.cu_seqlens.tolist() == np.concatenate([[0], np.cumsum(b.seqlens)]).tolist()
    sh = PackedLoader(ds, batch_size=4, shuffle=True, seed=3, device="cpu", drop_last=True)
    got = [b.ids for b in sh]
    assert len(got) == 1 and len(set(got[0])) == 4


def test_open_errors(tmp_path):
    bad = tmp_path / "bad.vspack"
    bad.write_bytes(b"not a pack file" * 10)
    with pytest.raises(_cabi.VsumError):
        PackedDataset(str(bad))
    with pytest.raises(_cabi.VsumError):
        PackedDataset(str(tmp_path / "missing.vspack"))
    vids = _videos((20,))
    path = tmp_path / "t.vspack"
    write_pack(str(path), vids)
    data = path.read_bytes()
    (tmp_path / "trunc.vspack").write_bytes(data[: len(data) // 2])
    with pytest.raises(_cabi.VsumError):
        PackedDataset(str(tmp_path / "trunc.vspack"))


def test_optional_arrays_and_pretrain_collate(tmp_path):
    """Videos without annotations (pretraining data, dataset.py:14-37): only features + video_rep; arrays that were not
    written come back as None and the pretrain collate stacks the video representations."""
    rng = np.random.default_rng(5)
    vids = [dict(name=f"p{i}", features=rng.random((n, 1024), dtype=np.float32), video_rep=rng.random(512, dtype=np.float32))
            for i, n in enumerate((40, 3, 77))]
    path = str(tmp_path / "p.vspack")
    write_pack(path, vids)
    ds = PackedDataset(path, split="pretrain")
    assert ds.array(0, _cabi.PACK_GTSCORE) is None and ds.array(1, _cabi.PACK_USER_SUMMARY) is None
    assert ds.array(2, _cabi.PACK_CHANGE_POINTS) is None and ds.array(0, _cabi.PACK_PICKS) is None
    (b,) = list(PackedLoader(ds, batch_size=8, device="cpu"))
    assert b.targets is None and b.video_rep.shape == (3, 512)
    assert np.array_equal(b.video_rep.numpy(), np.stack([v["video_rep"] for v in vids]))
    assert np.array_equal(b.features.numpy(), np.concatenate([v["features"] for v in vids])) and b.cu_seqlens.tolist() == [0, 40, 43, 120]
    with pytest.raises(_cabi.VsumError):                      # the train split needs gtscore, which this file does not have
        list(PackedLoader(PackedDataset(path, split="train"), batch_size=2, device="cpu"))
    with pytest.raises(ValueError):
        write_pack(str(tmp_path / "bad.vspack"), [dict(name="x", features=np.zeros((4, 8), np.float32))])


@pytest.mark.parametrize("u8", [False, True])
def test_eval_collate_matches_the_host_batch_builder(tmp_path, u8):
    """vsum_pack_eval_collate (native metadata gather of an evaluation batch) against HostEvalBatch.build over the same
    videos in pack_videos' longest-first order: every array, the knapsack order and the per-class launches."""
    import ctypes as C
    from vsum_b200.evaluation import _engine
    ns = (60, 7, 131, 300, 1, 2000, 131, 900)
    vids = _videos(ns, first=900)
    path = str(tmp_path / "d.vspack")
    write_pack(path, vids, user_summary_u8=u8)
    ds = PackedDataset(path, split="val")
    L = _cabi.load()
    sel = [5, 0, 2, 6, 7, 4, 1]                                             # a batch in arbitrary dataset order
    raw = np.asarray([ds.ids[i] for i in sel], dtype=np.int32)
    lay = _cabi.EvalBatchLayout()
    _cabi.check(L.vsum_pack_eval_collate(ds._h, raw.ctypes.data, len(sel), None, 0, C.byref(lay)), "layout")
    blob = np.zeros(int(lay.blob_bytes), np.uint8)
    lay2 = _cabi.EvalBatchLayout()
    _cabi.check(L.vsum_pack_eval_collate(ds._h, raw.ctypes.data, len(sel), blob.ctypes.data, blob.nbytes, C.byref(lay2)), "collate")
    assert bytes(lay) == bytes(lay2)
    order = sorted(range(len(sel)), key=lambda k: -ns[sel[k]])               # stable, longest first
    vs = [vids[sel[k]] for k in order]
    ref = _engine.HostEvalBatch.build([v["change_points"] for v in vs], [v["n_frames"] for v in vs], [v["picks"] for v in vs],
                                      [v["user_summary"].astype(np.uint8 if u8 else np.float32) for v in vs])
    B = len(sel)
    view = lambda off, n, dt: np.frombuffer(blob, dtype=dt, count=n, offset=int(off))
    assert view(lay.off_video_ids, B, np.int32).tolist() == [sel[k] for k in order]
    assert view(lay.off_cu_steps, B + 1, np.int32).tolist() == _engine._cu([len(v["picks"]) for v in vs]).tolist()
    assert np.array_equal(view(lay.off_picks, int(lay.T), np.int32), ref.picks)
    for name, dt, n in (("cu_picks", np.int32, B + 1), ("n_frames", np.int32, B), ("cu_shots", np.int32, B + 1),
                        ("bit_offsets", np.int64, B + 1), ("order", np.int32, B), ("sum_offsets", np.int64, B + 1),
                        ("us_offsets", np.int64, B + 1), ("cu_users", np.int32, B + 1), ("us_cols", np.int32, B)):
        assert np.array_equal(view(getattr(lay, "off_" + name), n, dt), getattr(ref, name)), name
    assert np.array_equal(view(lay.off_cps, 2 * lay.total_shots, np.int32).reshape(-1, 2), ref.cps)
    assert [(lay.launch_first[i], lay.launch_count[i], lay.launch_max_cap[i]) for i in range(lay.n_launches)] == ref.launches
    assert (lay.max_cap, lay.total_users, lay.us_elems, lay.summary_frames, lay.bit_words) == \
        (ref.max_cap, int(ref.cu_users[-1]), int(ref.us_offsets[-1]), int(ref.sum_offsets[-1]), int(ref.bit_offsets[-1]))
    assert lay.user_summary_dtype == (_cabi.USER_SUMMARY_U8 if u8 else _cabi.USER_SUMMARY_F32)
    # DMA out of the dataset needs page-locked residency: a mapped pack is refused, loudly
    assert L.vsum_pack_residency(ds._h) == _cabi.PACK_MMAP
    assert L.vsum_pack_h2d(ds._h, blob.ctypes.data, C.byref(lay), None, None, None) != 0
    assert b"VSUM_PACK_PINNED" in L.vsum_last_error()
    bad = np.asarray([99], np.int32)
    assert L.vsum_pack_eval_collate(ds._h, bad.ctypes.data, 1, None, 0, C.byref(lay)) != 0


def test_eval_collate_edge_cases(tmp_path):
    """Empty batch, a one-video batch and a pack without user summaries."""
    import ctypes as C
    vids = _videos((33, 300), first=950)
    for v in vids:
        del v["user_summary"], v["user_scores"]
    path = str(tmp_path / "d.vspack")
    write_pack(path, vids)
    ds = PackedDataset(path, split="val")
    L = _cabi.load()
    lay = _cabi.EvalBatchLayout()
    _cabi.check(L.vsum_pack_eval_collate(ds._h, None, 0, None, 0, C.byref(lay)), "empty")
    assert (lay.B, lay.T, lay.total_shots, lay.total_users, lay.n_launches) == (0, 0, 0, 0, 0)
    raw = np.asarray([1], np.int32)
    _cabi.check(L.vsum_pack_eval_collate(ds._h, raw.ctypes.data, 1, None, 0, C.byref(lay)), "one")
    blob = np.zeros(int(lay.blob_bytes), np.uint8)
    _cabi.check(L.vsum_pack_eval_collate(ds._h, raw.ctypes.data, 1, blob.ctypes.data, blob.nbytes, C.byref(lay)), "one")
    assert (lay.B, lay.T, lay.total_users, lay.us_elems, lay.n_launches) == (1, 300, 0, 0, 1)
    assert np.frombuffer(blob, np.int32, 2, int(lay.off_cu_steps)).tolist() == [0, 300]
    assert np.array_equal(np.frombuffer(blob, np.int32, 300, int(lay.off_picks)), vids[1]["picks"])
    small = np.zeros(16, np.uint8)                                     # a blob that is too small is refused
    assert L.vsum_pack_eval_collate(ds._h, raw.ctypes.data, 1, small.ctypes.data, small.nbytes, C.byref(lay)) != 0
