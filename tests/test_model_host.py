"""CPU: host-side logic of the model and evaluation packages (no kernels run)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import REFERENCE, ROOT
from vsum_b200.evaluation import _engine
from vsum_b200.model import PretrainModel, SimNet
from vsum_b200.sharding import partition, scorer_cost
from vsum_b200.synthetic import make_video

EXPECTED_KEYS_L1 = [
    "embedding_layer.feature_transform.weight", "embedding_layer.feature_transform.bias",
    "embedding_layer.positional_encoding.pos_embedding",
    "encoder.module_list.0.sa.q.weight", "encoder.module_list.0.sa.q.bias",
    "encoder.module_list.0.sa.k.weight", "encoder.module_list.0.sa.k.bias",
    "encoder.module_list.0.sa.v.weight", "encoder.module_list.0.sa.v.bias",
    "encoder.module_list.0.sa.feature_projection.weight", "encoder.module_list.0.sa.feature_projection.bias",
    "encoder.module_list.0.mlp.fc1.weight", "encoder.module_list.0.mlp.fc1.bias",
    "encoder.module_list.0.mlp.fc2.weight", "encoder.module_list.0.mlp.fc2.bias",
    "encoder.module_list.0.norm1.weight", "encoder.module_list.0.norm1.bias",
    "encoder.module_list.0.norm2.weight", "encoder.module_list.0.norm2.bias",
    "final_layer.weight", "final_layer.bias",
]


def test_state_dict_contract(seeded_model_kwargs):
    m = SimNet(num_heads=4, d_model=256, num_layers=1)
    assert list(m.state_dict().keys()) == EXPECTED_KEYS_L1
    full = SimNet(**seeded_model_kwargs)
    assert len(full.state_dict()) == 69                                  # SURVEY section 5 (checkpoint row)
    assert sum(p.numel() for p in full.parameters()) == 3_421_697
    assert full.state_dict()["embedding_layer.positional_encoding.pos_embedding"].shape == (1, 2000, 256)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference not mounted")
def test_seeded_init_is_bit_identical_to_reference(seeded_model_kwargs):
    import importlib
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] == "model"}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, REFERENCE)
    try:
        ref = importlib.import_module("model")
        torch.manual_seed(1234)
        r = ref.SimNet(**seeded_model_kwargs).state_dict()
    finally:
        sys.path.remove(REFERENCE)
        for k in [k for k in sys.modules if k.split(".")[0] == "model"]:
            del sys.modules[k]
        sys.modules.update(saved)
    torch.manual_seed(1234)
    mine = SimNet(**seeded_model_kwargs)
    sd = mine.state_dict()
    assert list(sd.keys()) == list(r.keys())
    for k in r:
        assert torch.equal(sd[k], r[k]), k
    mine.load_state_dict(r, strict=True)


def test_pretrain_wrapper_contract():
    m = PretrainModel(feature_dim=256, sparsity=0.0, num_heads=4, num_layers=2, dropout=0.2, use_pos=True)
    assert isinstance(m.encoder, SimNet) and m.video_transform.out_features == 512
    assert any(k.startswith("encoder.encoder.module_list.1.") for k in m.state_dict())


def test_dropin_package_names():
    import importlib
    import subprocess
    code = ("import sys; sys.path.insert(0, r'%s'); sys.path.insert(0, r'%s');"
            "from model import SimNet, PretrainModel;"
            "from evaluation.compute_metrics import eval_metrics;"
            "from evaluation.generate_summary import generate_summary;"
            "from evaluation.knapsack_implementation import knapSack;"
            "from evaluation.evaluation_metrics import evaluate_summary;"
            "from utils import set_seed, AverageMeter, load_json, load_yaml, mse_with_mask_loss;"
            "import vsum_b200.model.simnet as s; assert SimNet is s.SimNet; print('ok')"
            % (os.path.join(ROOT, "video-summarization_b200"), os.path.join(ROOT, "video-summarization_b200", "dropin")))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_host_eval_batch_packing():
    vids = [make_video(i, n, n_users=u, with_features=False) for i, (n, u) in enumerate([(5, 2), (300, 20), (40, 1)])]
    hb = _engine.HostEvalBatch.build([v.change_points for v in vids], [np.array(v.n_frames) for v in vids],
                                     [v.picks for v in vids], [v.user_summary for v in vids])
    assert hb.B == 3 and hb.cu_picks.tolist() == [0, 5, 305, 345]
    assert hb.cu_shots[-1] == sum(len(v.change_points) for v in vids)
    assert hb.sum_offsets.tolist() == np.cumsum([0] + [v.n_frames for v in vids]).tolist()
    assert hb.cu_users.tolist() == [0, 2, 22, 23]
    caps = [int(v.n_frames * 0.15) for v in vids]
    assert hb.max_cap == max(caps)
    widths = _engine.knapsack_class_width(np.array(caps))
    words = [len(v.change_points) * int(w // 32) for v, w in zip(vids, widths)]
    assert hb.bit_offsets.tolist() == np.cumsum([0] + words).tolist()
    assert hb.order[0] == 1                                              # largest capacity first
    assert sum(c for _, c, _ in hb.launches) == 3 and hb.launches[0] == (0, 1, caps[1])


def test_user_summary_dtype_selection_and_bf16_rounding():
    """Byte user summaries stay bytes only when EVERY video brings them as uint8 / bool (one dtype per batch); the pack
    writer's float32 -> bfloat16 rounding is round-to-nearest-even, i.e. torch's."""
    import torch
    from vsum_b200.data.packed import to_bf16_bits
    vids = [make_video(900 + i, n, n_users=3, with_features=False) for i, n in enumerate((20, 64))]
    build = lambda us: _engine.HostEvalBatch.build([v.change_points for v in vids], [np.array(v.n_frames) for v in vids],
                                                   [v.picks for v in vids], us)
    all_u8 = build([v.user_summary.astype(np.uint8) for v in vids])
    assert all_u8.user_summary.dtype == np.uint8 and _engine._us_dtype_code(all_u8.user_summary) == 1
    assert build([v.user_summary.astype(bool) for v in vids]).user_summary.dtype == np.uint8
    mixed = build([vids[0].user_summary.astype(np.uint8), vids[1].user_summary])
    assert mixed.user_summary.dtype == np.float32 and _engine._us_dtype_code(mixed.user_summary) == 0
    assert np.array_equal(mixed.user_summary, all_u8.user_summary.astype(np.float32))
    assert all_u8.us_offsets.tolist() == mixed.us_offsets.tolist()
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.standard_normal(4096).astype(np.float32) * 10.0 ** rng.integers(-20, 20, 4096).astype(np.float32),
                        np.array([0.0, -0.0, 1.00390625, 1.01171875, 3.3895314e38, -1e-40, np.inf], np.float32)])
    want = torch.from_numpy(x).bfloat16().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(to_bf16_bits(x), want)


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(3)
    ns = [int(x) for x in np.exp(rng.uniform(np.log(128), np.log(8192), 500))]
    for world in (1, 2, 4, 8):
        shards = partition(ns, world)
        assert sorted(i for s in shards for i in s) == list(range(len(ns)))
        loads = [sum(scorer_cost(ns[i]) for i in s) for s in shards]
        assert max(loads) / (sum(loads) / world) < 1.05


def test_masked_mse_has_no_cpu_path_and_oracle_keeps_reference_semantics():
    from oracle import scorer_ref
    from vsum_b200._cabi import VsumError
    from vsum_b200.utils import mse_with_mask_loss
    out = torch.tensor([[[1.0], [2.0], [9.0]]])
    tgt = torch.tensor([[0.0, 0.0, 1000.0]])
    mask = torch.tensor([[False, False, True]])
    assert scorer_ref.masked_mse(out, tgt, mask).item() == pytest.approx((1 + 4 + 0) / 3)   # divides by bs*Nmax (utils.py:55)
    with pytest.raises(VsumError):
        mse_with_mask_loss(out, tgt, mask)
