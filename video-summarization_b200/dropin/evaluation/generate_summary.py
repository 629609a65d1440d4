"""Shim: `evaluation.generate_summary` of the reference -> `vsum_b200.evaluation.generate_summary`."""
from vsum_b200.evaluation.generate_summary import *  # noqa: F401,F403
from vsum_b200.evaluation import generate_summary as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
