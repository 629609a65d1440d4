"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

CPU restatement (pure Python / numpy, same execution model as the reference) of the evaluation
half of the hot path of BerserkerMother/Video-Summarization:

    scores -> upsample -> per-shot float32 mean -> 15 % capacity -> 0/1 knapsack
           -> int8 summary mask -> per-user overlap P/R/F -> 'avg' | 'max'

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import this file.  Each function cites the reference file:line it restates (paths relative to
the reference root).

Pinning: the reference ships no tests or golden vectors; its only known
answer is the commented driver at `src/evaluation/knapsack_implementation.py:35-41`
(-> [0, 1, 2, 3, 4]).  Parity is therefore pinned by (1) that known answer, (2) fixtures under
`tests/golden/` produced by importing the reference itself in the build container
(`tests/golden/make_golden.py`, committed next to them) and (3) a live differential test that runs
whenever `/root/reference` is present (`tests/test_oracle_vs_reference.py`).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# numpy float32 pairwise summation (third-party arithmetic: numpy 2.3.5 `pairwise_sum`,
# reached from `src/evaluation/generate_summary.py:42` through `ndarray.mean`).
# SURVEY.md Appendix A.1 states the algorithm; restated here with explicit float32 rounding.
# --------------------------------------------------------------------------------------
def pairwise_sum_f32(a: np.ndarray, lo: int, n: int) -> np.float32:
    f32 = np.float32
    if n < 8:
        # numpy starts from -0.0 (float32) so an all-(-0.0) slice keeps its sign; for any
        # other input the result equals starting from a[lo].
        res = f32(-0.0)
        for i in range(n):
            res = f32(res + a[lo + i])
        return res
    if n <= 128:
        r = [f32(a[lo + j]) for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = f32(r[j] + a[lo + i + j])
            i += 8
        res = f32(f32(f32(r[0] + r[1]) + f32(r[2] + r[3])) + f32(f32(r[4] + r[5]) + f32(r[6] + r[7])))
        while i < n:
            res = f32(res + a[lo + i])
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return f32(pairwise_sum_f32(a, lo, n2) + pairwise_sum_f32(a, lo + n2, n - n2))


def mean_f32(a: np.ndarray) -> float:
    """`a.mean().item()` for a contiguous float32 slice: float32 sum / float32 count -> fp64."""
    n = len(a)
    if n == 0:
        return float("nan")
    total = np.float32(np.float32(0.0) + pairwise_sum_f32(a, 0, n))   # add.reduce starts from +0.0
    return float(np.float32(total / np.float32(n)))


# --------------------------------------------------------------------------------------
# upsample  (src/evaluation/generate_summary.py:25-35 == src/evaluation/compute_metrics.py:19-39)
# --------------------------------------------------------------------------------------
def upsample_scores(scores: np.ndarray, n_frames: int, picks: np.ndarray) -> np.ndarray:
    out = np.zeros(int(n_frames), dtype=np.float32)
    pos = np.asarray(picks)
    if pos.dtype != int:
        pos = pos.astype(np.int32)
    if pos[-1] != n_frames:
        pos = np.concatenate([pos, [n_frames]])
    for i in range(len(pos) - 1):
        out[pos[i]:pos[i + 1]] = 0 if i == len(scores) else scores[i]
    return out


# --------------------------------------------------------------------------------------
# 0/1 knapsack  (src/evaluation/knapsack_implementation.py:11-28)
# --------------------------------------------------------------------------------------
def knapsack_select(capacity: int, weights, values, n: int) -> list[int]:
    prev = [0] * (capacity + 1)
    take_rows = []
    for i in range(1, n + 1):
        wt, val = weights[i - 1], values[i - 1]
        cur = [0] * (capacity + 1)
        took = [False] * (capacity + 1)
        for w in range(1, capacity + 1):
            if wt <= w:
                a, b = val + prev[w - wt], prev[w]
                cur[w] = max(a, b)          # Python max: keeps `a` unless b > a
            else:
                cur[w] = prev[w]
            took[w] = cur[w] != prev[w]     # the reference's back-track test, line 26
        take_rows.append(took)
        prev = cur
    chosen, w = [], capacity
    for i in range(n, 0, -1):
        if take_rows[i - 1][w]:
            chosen.insert(0, i - 1)
            w -= weights[i - 1]
    return chosen


def capacity_of(last_shot_end: int) -> int:
    """src/evaluation/generate_summary.py:45-46 -- fp64 multiply then truncation."""
    return int((last_shot_end + 1) * 0.15)


# --------------------------------------------------------------------------------------
# generate_summary for one video  (src/evaluation/generate_summary.py:17-55)
# --------------------------------------------------------------------------------------
def summarize_video(change_points, scores, n_frames, picks):
    """Returns (summary int8[last_end+1], shot_means fp64[S], shot_lengths int[S], capacity, selected)."""
    frame_scores = upsample_scores(scores, int(n_frames), picks)
    lengths, means = [], []
    for s, e in change_points:
        lengths.append(int(e) - int(s) + 1)
        means.append(mean_f32(frame_scores[int(s):int(e) + 1]))
    last_end = int(change_points[-1][1])
    cap = capacity_of(last_end)
    selected = knapsack_select(cap, lengths, means, len(lengths))
    summary = np.zeros(last_end + 1, dtype=np.int8)
    for k in selected:
        summary[int(change_points[k][0]):int(change_points[k][1]) + 1] = 1
    return summary, np.array(means, dtype=np.float64), np.array(lengths, dtype=np.int64), cap, selected


# --------------------------------------------------------------------------------------
# keyshot F-score  (src/evaluation/evaluation_metrics.py:12-33)
# --------------------------------------------------------------------------------------
def fscore_users(summary: np.ndarray, user_summary: np.ndarray, builtin_sums: bool = False) -> list:
    """Per-user F (x100).  Counts are integers; ratios are IEEE fp64 with 0/0 -> NaN.
    `builtin_sums`: count with Python's builtin `sum()` over the numpy arrays, element by element, exactly as
    evaluation_metrics.py:23-24 does (same integers, ~100x the time: the reference's slowest F-score loop) -- the
    timing arm of bench.py uses it; the default `ndarray.sum()` keeps the test suite fast."""
    width = max(len(summary), user_summary.shape[1])
    S = np.zeros(width, dtype=np.int64)
    S[:len(summary)] = summary
    count = (lambda a: int(sum(a))) if builtin_sums else (lambda a: int(a.sum()))
    out = []
    with np.errstate(divide="ignore", invalid="ignore"):
        for u in range(user_summary.shape[0]):
            G = np.zeros(width, dtype=np.int64)
            G[:user_summary.shape[1]] = user_summary[u]          # float -> int truncation
            s_cnt = count(S)                                      # the reference re-counts S for every user (line 23)
            o_cnt, g_cnt = count(S & G), count(G)
            p = np.float64(o_cnt) / np.float64(s_cnt)
            r = np.float64(o_cnt) / np.float64(g_cnt)
            out.append(0 if p + r == 0 else 2 * p * r * 100 / (p + r))
    return out


def fscore_reduce(per_user: list, method: str):
    if method == "max":
        return max(per_user)
    return sum(per_user) / len(per_user)


def fscore_video(summary, user_summary, method: str = "avg", builtin_sums: bool = False):
    return fscore_reduce(fscore_users(summary, user_summary, builtin_sums), method)


def pool_eval_video(args):
    """multiprocessing worker of bench.py's `fair` CPU figure: the pure-Python stages of one video
    (generate_summary + evaluate_summary as shipped, builtin sums included) given its scores."""
    change_points, scores, n_frames, picks, user_summary, method = args
    summary, *_ = summarize_video(change_points, scores, n_frames, picks)
    return fscore_video(summary, user_summary, method, builtin_sums=True)


# --------------------------------------------------------------------------------------
# eval_metrics minus the scipy correlations  (src/evaluation/compute_metrics.py:42-92)
# --------------------------------------------------------------------------------------
def evaluate_videos(score_dict: dict, user_dict: dict, method: str = "avg"):
    """Returns (mean F, per-video F list) in dict insertion order."""
    per_video = []
    for name, scores in score_dict.items():
        u = user_dict[name]
        summary, *_ = summarize_video(u.change_points, scores, int(u.n_frames), u.picks)
        per_video.append(fscore_video(summary, u.user_summary, method))
    return np.mean(per_video), per_video
