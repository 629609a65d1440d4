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
