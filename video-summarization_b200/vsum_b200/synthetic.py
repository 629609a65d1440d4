"""Deterministic synthetic videos shaped like the DSNet TVSum/SumMe h5 records.

There is no network for datasets, so every test and bench line runs on these.  A video is a
function of its integer id only (`np.random.default_rng(1234 + v)`), so the same video is
produced on every rank / in every process regardless of how the set is sharded
(SURVEY.md §8(d)).  Field names follow the reference's `UserSummaries` record
(reference `src/data/dataset.py:146-154`) and its h5 reads (`dataset.py:93-103`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

IN_FEATURES = 1024          # reference `src/model/simnet.py:22`
PICK_STRIDE = 15            # DSNet h5 `picks` keep every 15th original frame
FRAMES_PER_SHOT = 150       # mean shot length used to size the change-point list
PAD_SENTINEL = 1000.0       # reference `src/train.py:118`, `src/data/dataset.py:159`


class UserSummaries:
    """Mirror of the reference record (`src/data/dataset.py:146-154`), same attribute names."""

    def __init__(self, user_summary, user_scores, name, changes_point, n_frames, picks):
        self.user_summary = user_summary
        self.user_scores = user_scores
        self.change_points = changes_point
        self.n_frames = n_frames
        self.picks = picks
        self.name = name


@dataclass
class SyntheticVideo:
    vid: int
    name: str
    n_steps: int                      # N, sub-sampled frames
    n_frames: int                     # original frames
    picks: np.ndarray                 # int32[N]
    change_points: np.ndarray         # int32[S,2], inclusive [start,end]
    user_summary: np.ndarray          # float32[U,n_frames] of 0/1
    features: Optional[np.ndarray]    # float32[N,1024] or None
    gtscore: Optional[np.ndarray]     # float32[N] or None
    user_scores: Optional[np.ndarray]  # float32[U,n_frames] or None

    def as_user(self) -> UserSummaries:
        return UserSummaries(self.user_summary, self.user_scores, self.name,
                             self.change_points, np.array(self.n_frames), self.picks)


def video_length(v: int, lo: int, hi: int, seed: int = 99) -> int:
    """Log-uniform N in [lo, hi] for video id `v` (config 5: N=128..8192)."""
    rng = np.random.default_rng(seed * 1_000_003 + v)
    if lo == hi:
        return int(lo)
    return int(round(float(np.exp(rng.uniform(np.log(lo), np.log(hi))))))


def make_video(v: int, n_steps: int, n_users: int = 20, with_features: bool = True,
               with_user_scores: bool = False) -> SyntheticVideo:
    """Build video `v` with `n_steps` sub-sampled frames.  Draw order is part of the recipe:
    n_frames jitter, cuts, per-user shot marks, forced shot, features, gtscore, user_scores."""
    assert n_steps >= 1
    rng = np.random.default_rng(1234 + v)
    n_frames = PICK_STRIDE * (n_steps - 1) + 1 + int(rng.integers(0, PICK_STRIDE))
    picks = np.arange(0, n_frames, PICK_STRIDE, dtype=np.int32)
    assert len(picks) == n_steps
    n_shots = max(2, int(round(n_frames / FRAMES_PER_SHOT)))
    n_shots = min(n_shots, n_frames)          # tiny videos: at most one shot per frame
    if n_shots >= 2:
        cuts = np.sort(rng.choice(np.arange(1, n_frames), n_shots - 1, replace=False))
    else:
        cuts = np.zeros(0, dtype=np.int64)
    starts = np.concatenate([[0], cuts]).astype(np.int32)
    ends = np.concatenate([cuts - 1, [n_frames - 1]]).astype(np.int32)
    change_points = np.stack([starts, ends], axis=1)

    marks = rng.random((n_users, n_shots)) < 0.15
    forced = rng.integers(0, n_shots, size=n_users)
    marks[np.arange(n_users), forced] = True          # reference returns NaN for an empty user row
    shot_len = (ends - starts + 1).astype(np.int64)
    user_summary = np.repeat(marks, shot_len, axis=1).astype(np.float32)

    features = gtscore = user_scores = None
    if with_features:
        features = rng.random((n_steps, IN_FEATURES), dtype=np.float32)
        gtscore = rng.random(n_steps, dtype=np.float32)
    if with_user_scores:
        user_scores = rng.random((n_users, n_frames), dtype=np.float32)
    return SyntheticVideo(v, f"video_{v}", n_steps, n_frames, picks, change_points,
                          user_summary, features, gtscore, user_scores)


def make_scores(v: int, n_steps: int) -> np.ndarray:
    """Stand-in importance scores in (0,1) for evaluation-only tests (float32[N])."""
    rng = np.random.default_rng(777_000 + v)
    return rng.random(n_steps, dtype=np.float32) * np.float32(0.98) + np.float32(0.01)
