"""Shim: `evaluation.evaluation_metrics` of the reference -> `vsum_b200.evaluation.evaluation_metrics`."""
from vsum_b200.evaluation.evaluation_metrics import *  # noqa: F401,F403
from vsum_b200.evaluation import evaluation_metrics as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
