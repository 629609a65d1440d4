"""vsum_b200 -- B200-native drop-in for the video-summarisation hot path of
BerserkerMother/Video-Summarization: `model.SimNet` (frame scorer) and `evaluation.*`
(shot pooling, knapsack, keyshot F-score), executed by hand-written sm_100a kernels behind the
C ABI declared in `include/vsum_b200.h`."""
from . import _cabi  # noqa: F401
from .synthetic import UserSummaries  # noqa: F401

__all__ = ["model", "evaluation", "utils", "pipeline", "synthetic", "sharding"]
