"""Drop-in for the reference's `src/utils` package (train.py:15)."""
from vsum_b200.utils import set_seed, AverageMeter, load_yaml, load_json, mse_with_mask_loss  # noqa: F401
