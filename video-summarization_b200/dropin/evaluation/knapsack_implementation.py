"""Shim: `evaluation.knapsack_implementation` of the reference -> `vsum_b200.evaluation.knapsack_implementation`."""
from vsum_b200.evaluation.knapsack_implementation import *  # noqa: F401,F403
from vsum_b200.evaluation import knapsack_implementation as _impl

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
