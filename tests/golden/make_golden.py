"""Generates the golden fixtures in this directory by running the REFERENCE ITSELF
(`/root/reference/src`, imported unmodified, CPU, fp32) on seeded synthetic inputs.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Inputs are never stored: they are regenerated from `vsum_b200.synthetic` (numpy Generator seeded
with the video id) and `torch.manual_seed(1234)` weights, both deterministic for the pinned
numpy 2.3 / torch 2.11 of this image.  Versions are recorded in each fixture.
"""
from __future__ import annotations

import importlib
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/src"
sys.path.insert(0, os.path.join(ROOT, "video-summarization_b200"))

from vsum_b200.synthetic import make_scores, make_video  # noqa: E402

# (video id, N, users): edge cases first -- single frame, fewer frames than a shot, SumMe/TVSum
# sized, the reference's 2000-row positional limit, long videos for the evaluation path.
EVAL_CASES = [(0, 1, 1), (1, 2, 3), (2, 7, 15), (3, 11, 20), (4, 60, 15), (5, 170, 20), (6, 300, 20),
              (7, 301, 20), (8, 650, 15), (9, 1024, 20), (10, 1300, 20), (11, 2000, 20), (12, 2047, 5),
              (13, 4096, 20), (14, 8192, 20), (15, 333, 1)]
SCORER_CASES = [(100, 1), (101, 5), (102, 130), (103, 300), (104, 777), (105, 2000)]
SCORER_LONG = (106, 2300)          # needs the longer positional table (reference class, maxlen=4096)
PADDED_LENS = (300, 180, 77)       # train.py:118 style padded batch, ids 110..112
MODEL_KW = dict(num_heads=4, d_model=256, num_layers=4, sparsity=0., use_cls=False, dropout=0.3,
                num_classes=1, use_pos=True)            # run_finetune.sh:1 / train.py:29-34
SMALL_KW = dict(num_heads=4, d_model=64, num_layers=2, sparsity=0., use_cls=False, dropout=0.1,
                num_classes=1, use_pos=True)


def import_reference():
    assert os.path.isdir(REF), "the reference is only mounted in the build container"
    sys.path.insert(0, REF)
    mods = {name: importlib.import_module(name) for name in ("model", "evaluation", "utils")}
    assert mods["model"].__file__.startswith(REF)
    return mods


def versions():
    return np.array([f"torch {torch.__version__}", f"numpy {np.__version__}"])


def eval_goldens(ev):
    from evaluation.generate_summary import generate_summary
    from evaluation.evaluation_metrics import evaluate_summary
    out = {"versions": versions(), "cases": np.array(EVAL_CASES, dtype=np.int64)}
    for vid, n, users in EVAL_CASES:
        v = make_video(vid, n, n_users=users, with_features=False)
        sc = make_scores(vid, n)
        summary = generate_summary([v.change_points], [sc], [np.array(v.n_frames)], [v.picks])[0]
        # shot means exactly as generate_summary.py:38-42 computes them
        from evaluation.compute_metrics import upsample
        fs = upsample(sc, v.n_frames, v.picks)
        means = np.array([fs[s:e + 1].mean().item() for s, e in v.change_points], dtype=np.float64)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f_avg = evaluate_summary(summary, v.user_summary, "avg")
            f_max = evaluate_summary(summary, v.user_summary, "max")
        out[f"summary_{vid}"] = np.packbits(summary.astype(np.uint8))
        out[f"means_{vid}"] = means
        out[f"f_{vid}"] = np.array([f_avg, f_max], dtype=np.float64)
    # known answer (knapsack_implementation.py:35-41) and tie-breaking cases
    from evaluation.knapsack_implementation import knapSack
    out["knap_driver"] = np.array(knapSack(7, [2, 2, 1, 1, 1, 2], [4, 4, 2, 2, 2, 4], 6))
    out["knap_tie_a"] = np.array(knapSack(2, [2, 1, 1], [2, 1, 1], 3))
    out["knap_tie_b"] = np.array(knapSack(2, [1, 1, 2], [1, 1, 2], 3))
    np.savez_compressed(os.path.join(HERE, "eval_golden.npz"), **out)


def eval_metrics_golden(ev):
    """Whole `eval_metrics` (F + scipy correlations) on a 5-video split-sized set."""
    from evaluation.compute_metrics import eval_metrics
    data, users = {}, {}
    for vid, n in [(200, 210), (201, 340), (202, 275), (203, 398), (204, 305)]:
        v = make_video(vid, n, n_users=20, with_features=False, with_user_scores=True)
        data[v.name] = make_scores(vid, n)
        users[v.name] = v.as_user()
    f, tau, rho = eval_metrics(data, users)
    np.savez_compressed(os.path.join(HERE, "eval_metrics_golden.npz"), versions=versions(),
                        ids=np.array([200, 201, 202, 203, 204]), ns=np.array([210, 340, 275, 398, 305]),
                        result=np.array([f, tau, rho], dtype=np.float64))


@torch.no_grad()
def scorer_goldens(mods):
    torch.set_num_threads(8)
    mods["utils"].set_seed(1234)
    net = mods["model"].SimNet(**MODEL_KW).eval()
    out = {"versions": versions(), "cases": np.array(SCORER_CASES, dtype=np.int64)}
    for vid, n in SCORER_CASES:
        x = torch.from_numpy(make_video(vid, n).features).unsqueeze(0)
        logits, feats = net(x)
        out[f"logits_{vid}"] = logits.view(-1).numpy()
        out[f"feats_head_{vid}"] = feats[0, :4].numpy()
        out[f"feats_tail_{vid}"] = feats[0, -4:].numpy()
        out[f"feats_rowsum_{vid}"] = feats[0].double().sum(dim=1).numpy()
    # N > 2000: the reference's own PositionalEncoding class with a longer table
    simnet_mod = importlib.import_module("model.simnet")
    net.embedding_layer.positional_encoding = simnet_mod.PositionalEncoding(emb_size=256, dropout=0., maxlen=4096).eval()
    vid, n = SCORER_LONG
    logits, feats = net(torch.from_numpy(make_video(vid, n).features).unsqueeze(0))
    out[f"logits_{vid}"] = logits.view(-1).numpy()
    out[f"feats_rowsum_{vid}"] = feats[0].double().sum(dim=1).numpy()
    # padded batch + key mask (train.py:115-121), valid rows only
    mods["utils"].set_seed(1234)
    net = mods["model"].SimNet(**MODEL_KW).eval()
    nmax = max(PADDED_LENS)
    x = torch.full((len(PADDED_LENS), nmax, 1024), 1000.0)
    tgt = torch.full((len(PADDED_LENS), nmax), 1000.0)
    for b, n in enumerate(PADDED_LENS):
        v = make_video(110 + b, n)
        x[b, :n] = torch.from_numpy(v.features)
        tgt[b, :n] = torch.from_numpy(v.gtscore)
    mask = x[:, :, 0] == 1000
    logits, _ = net(x, mask)
    for b, n in enumerate(PADDED_LENS):
        out[f"padded_logits_{b}"] = logits[b, :n, 0].numpy()
    out["padded_loss"] = np.array(mods["utils"].mse_with_mask_loss(logits, tgt, mask).item(), dtype=np.float64)
    # small generic model for the fp32 kernels (d_model 64, 2 layers)
    mods["utils"].set_seed(1234)
    small = mods["model"].SimNet(**SMALL_KW).eval()
    logits, feats = small(torch.from_numpy(make_video(120, 50).features).unsqueeze(0))
    out["small_logits"] = logits.view(-1).numpy()
    out["small_feats"] = feats[0].numpy()
    np.savez_compressed(os.path.join(HERE, "scorer_golden.npz"), **out)


SCORER_XLONG = [(107, 4096), (108, 8192)]   # BASELINE config 5 goes to N = 8192


@torch.no_grad()
def scorer_long_goldens(mods):
    """Reference logits at N = 4096 and 8192: the reference's own SimNet with its own PositionalEncoding class built with
    maxlen=8192 (simnet.py:220-238; the shipped table stops at 2000 rows, simnet.py:188).  ~10 s of CPU each."""
    torch.set_num_threads(8)
    mods["utils"].set_seed(1234)
    net = mods["model"].SimNet(**MODEL_KW).eval()
    simnet_mod = importlib.import_module("model.simnet")
    net.embedding_layer.positional_encoding = simnet_mod.PositionalEncoding(emb_size=256, dropout=0., maxlen=8192).eval()
    out = {"versions": versions(), "cases": np.array(SCORER_XLONG, dtype=np.int64)}
    for vid, n in SCORER_XLONG:
        logits, feats = net(torch.from_numpy(make_video(vid, n).features).unsqueeze(0))
        out[f"logits_{vid}"] = logits.view(-1).numpy()
        out[f"feats_rowsum_{vid}"] = feats[0].double().sum(dim=1).numpy()
    np.savez_compressed(os.path.join(HERE, "scorer_long_golden.npz"), **out)
    print("scorer_long_golden.npz written")


KTS_CASES = [(0, 80, 32, 10, 1, 100000), (1, 150, 64, 20, 1, 100000), (2, 60, 16, 5, 3, 20), (3, 40, 8, 39, 1, 100000),
             (4, 33, 4, 0, 1, 100000), (5, 300, 128, 40, 2, 100000)]       # (seed, n, dim, ncp, lmin, lmax)


def kts_features(seed, n, dim):
    """Frame features with a few plateaus (so change points exist); float32 like the h5 features."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim), dtype=np.float32)
    for c in np.sort(rng.choice(np.arange(1, n), size=min(6, n - 1), replace=False)):
        x[c:] += rng.random(dim, dtype=np.float32)
    return x


def kts_golden():
    """Runs the reference's kts_segmentation (src/data/preprocess/segmentations/kts) on seeded kernels."""
    import contextlib, importlib.util, io
    d = os.path.join(REF, "data/preprocess/segmentations/kts")          # loaded as its own package: `data` needs h5py
    spec = importlib.util.spec_from_file_location("ref_kts", os.path.join(d, "__init__.py"), submodule_search_locations=[d])
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_kts"] = ref
    spec.loader.exec_module(ref)
    out = {"versions": versions(), "cases": np.array(KTS_CASES, dtype=np.int64)}
    for seed, n, dim, ncp, lmin, lmax in KTS_CASES:
        x = kts_features(seed, n, dim)
        K = np.dot(x, x.T)
        with contextlib.redirect_stdout(io.StringIO()):
            cps, costs = ref.kts_segmentation(K, ncp, 1.0, lmin=lmin, lmax=lmax)
            cps_fixed, scores = ref.cpd_nonlin(K, ncp, lmin=lmin, lmax=lmax, verbose=False)
        out[f"K_{seed}"], out[f"cps_{seed}"], out[f"costs_{seed}"] = K, cps, costs
        out[f"cpsfixed_{seed}"], out[f"scores_{seed}"] = cps_fixed, scores
    np.savez_compressed(os.path.join(HERE, "kts_golden.npz"), **out)
    print("kts_golden.npz written")


@torch.no_grad()
def pretrain_golden(mods):
    sp = importlib.import_module("model.simnet_pretrain")
    sp.device = torch.device("cpu")                     # simnet_pretrain.py:9,64 uses a module global
    mods["utils"].set_seed(1234)
    net = mods["model"].PretrainModel(feature_dim=256, sparsity=0.0, num_heads=4, num_layers=4, dropout=0.2,
                                      use_pos=True).eval()
    lens = (120, 77)
    x = torch.full((2, 120, 1024), 1000.0)
    for b, n in enumerate(lens):
        x[b, :n] = torch.from_numpy(make_video(130 + b, n).features)
    rng = np.random.default_rng(4321)
    vid_rep = torch.from_numpy(rng.random((2, 512), dtype=np.float32))
    mask = x[:, :, 0] == 1000
    loss, center, repel = net(x, vid_rep, mask)
    np.savez_compressed(os.path.join(HERE, "pretrain_golden.npz"), versions=versions(),
                        losses=np.array([loss.item(), center.item(), repel.item()], dtype=np.float64))


if __name__ == "__main__":
    mods = import_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "scorer_long":      # only the N = 4096 / 8192 scorer fixture
        scorer_long_goldens(mods)
        sys.exit(0)
    scorer_long_goldens(mods)
    eval_goldens(mods["evaluation"])
    eval_metrics_golden(mods["evaluation"])
    scorer_goldens(mods)
    pretrain_golden(mods)
    kts_golden()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")
