"""Host helpers with the reference's names (`src/utils/utils.py`).  Only `mse_with_mask_loss`
(lines 45-56) touches the hot path; in round 1 it is thin PyTorch glue over tensors the CUDA
scorer produced (the native fwd/bwd loss kernel is SURVEY.md section 8 row a9, scheduled with
the backward pass)."""
from __future__ import annotations

import json
import random

import numpy as np
import torch
import yaml


def set_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


class AverageMeter:
    def __init__(self):
        self.val, self.num = 0, 0

    def update(self, val, num):
        self.val += val
        self.num += num

    def avg(self):
        return self.val / self.num


def load_yaml(path):
    with open(path, "r") as f:
        return yaml.safe_load(f)


def load_json(path):
    with open(path) as f:
        return json.load(f)


def mse_with_mask_loss(output, targets, mask, reduction="avg"):
    """Masked MSE normalised by the PADDED size bs*Nmax, not by the valid frames (utils.py:55)."""
    keep = (~mask).to(output.dtype)
    err = ((output.squeeze(2) - targets) * keep) ** 2
    return err.mean() if reduction == "avg" else err.sum()
