"""Shim: `model.simnet` of the reference -> `vsum_b200.model.simnet`."""
from vsum_b200.model.simnet import SimNet  # noqa: F401
