"""Shim: `utils.utils` of the reference -> `vsum_b200.utils.utils`."""
from vsum_b200.utils.utils import set_seed, AverageMeter, load_yaml, load_json, mse_with_mask_loss  # noqa: F401
