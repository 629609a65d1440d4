"""`evaluate_summary` with the reference signature (`src/evaluation/evaluation_metrics.py:4-33`):
per-user overlap counts on the GPU (integer), precision / recall / F in fp64 in the reference's
evaluation order, 'max' (SumMe) or average (TVSum) over users."""
from __future__ import annotations

import numpy as np

from . import _engine


def evaluate_summary(predicted_summary, user_summary, eval_method):
    f = _engine.fscore_of_masks([np.asarray(predicted_summary)], [np.asarray(user_summary)],
                                "max" if eval_method == "max" else "avg")
    return f[0]
