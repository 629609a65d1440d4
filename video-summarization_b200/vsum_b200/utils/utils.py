"""Host helpers with the reference's names (`src/utils/utils.py`).  Only `mse_with_mask_loss`
(lines 45-56) touches the hot path: it runs the native fwd+bwd kernel behind `vsum_masked_mse`
(SURVEY.md section 8 row a9).  There is no CPU fallback: host tensors raise."""
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


class _MaskedMSE(torch.autograd.Function):
    """vsum_masked_mse: loss and d(loss)/d(output) in one kernel pass."""

    @staticmethod
    def forward(ctx, output, targets, mask, denom):
        from .. import _cabi
        out = output.contiguous().float()
        tgt = targets.contiguous().float()
        pad = mask.contiguous().to(torch.uint8)
        loss = torch.zeros((), dtype=torch.float32, device=out.device)
        d_out = torch.empty_like(out)
        with torch.cuda.device(out.device):
            _cabi.check(_cabi.load().vsum_masked_mse(out.data_ptr(), tgt.data_ptr(), pad.data_ptr(), out.numel(), float(denom),
                                                     loss.data_ptr(), 1.0, d_out.data_ptr(),
                                                     torch.cuda.current_stream(out.device).cuda_stream), "vsum_masked_mse")
        ctx.save_for_backward(d_out)
        ctx.shape = output.shape
        return loss

    @staticmethod
    def backward(ctx, g):
        (d_out,) = ctx.saved_tensors
        return (d_out * g).view(ctx.shape), None, None, None


def mse_with_mask_loss(output, targets, mask, reduction="avg", denom=None):
    """Masked MSE normalised by the PADDED size bs*Nmax, not by the valid frames (utils.py:55): the native kernel behind
    `vsum_masked_mse` (forward and backward).  CUDA tensors only -- this package has no CPU path."""
    if not output.is_cuda:
        from .._cabi import VsumError
        raise VsumError("mse_with_mask_loss: vsum_b200 runs on CUDA tensors only (no CPU fallback)")
    squeezed = output.squeeze(2)
    if denom is None:       # data-parallel callers pass sharding.global_loss_denominator(...)
        denom = float(squeezed.numel()) if reduction == "avg" else 1.0
    return _MaskedMSE.apply(squeezed, targets, mask, float(denom))
