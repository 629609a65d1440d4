"""Shim: `evaluation.compute_correlation` of the reference -> `vsum_b200.evaluation.compute_correlation`."""
from vsum_b200.evaluation.compute_correlation import *  # noqa: F401,F403
from vsum_b200.evaluation import compute_correlation as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
