"""GPU box: standalone timing of the evaluation kernels on the bench batch (no scorer running)."""
import sys, numpy as np, torch
sys.path.insert(0, "video-summarization_b200")
from vsum_b200 import _cabi
from vsum_b200.evaluation import _engine
from vsum_b200.synthetic import make_video, make_scores, video_length
V = 256
vids = [make_video(v, video_length(v, 128, 8192), n_users=20, with_features=False) for v in range(V)]
hb = _engine.HostEvalBatch.build([v.change_points for v in vids], [np.array(v.n_frames) for v in vids],
                                 [v.picks for v in vids], [v.user_summary for v in vids])
db = _engine.DeviceEvalBatch(hb)
scores = torch.from_numpy(np.concatenate([make_scores(v.vid, v.n_steps) for v in vids])).cuda()
cu = torch.from_numpy(_engine._cu([v.n_steps for v in vids]).astype(np.int32)).cuda()
for _ in range(3): _engine.summarize(db, scores, cu)
torch.cuda.synchronize()
_cabi.profile_begin()
for _ in range(10): _engine.summarize(db, scores, cu)
torch.cuda.synchronize()
prof = _cabi.profile_end()
print({k: round(v[0] / 10, 4) for k, v in prof.items() if v[1]}, "launches", hb.launches, "bits MB", hb.bit_offsets[-1] * 4 / 1e6)
# the same batch with the user summaries as bytes (user_summary_u8 packs): overlap_u8_kernel
hb8 = _engine.HostEvalBatch.build([v.change_points for v in vids], [np.array(v.n_frames) for v in vids],
                                  [v.picks for v in vids], [v.user_summary.astype(np.uint8) for v in vids])
db8 = _engine.DeviceEvalBatch(hb8)
for _ in range(3): _engine.summarize(db8, scores, cu)
torch.cuda.synchronize()
_cabi.profile_begin()
for _ in range(10): _engine.summarize(db8, scores, cu)
torch.cuda.synchronize()
prof8 = _cabi.profile_end()
print("uint8 user summaries:", {k: round(v[0] / 10, 4) for k, v in prof8.items() if v[1]},
      "user bytes MB", hb8.user_summary.nbytes / 1e6, "vs", hb.user_summary.nbytes / 1e6)
