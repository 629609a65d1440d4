"""Shim: `model.simnet_pretrain` of the reference -> `vsum_b200.model.simnet_pretrain`."""
from vsum_b200.model.simnet_pretrain import PretrainModel  # noqa: F401
