"""`PretrainModel` with the reference's constructor / attributes / forward contract
(`src/model/simnet_pretrain.py:12-100`): `.encoder` is the CUDA `SimNet`, `.video_transform`
a `Linear(feature_dim, 512)`, `forward(x, video_representation, mask)` returns
`(loss, center_loss, repel_loss)`.

Everything runs on the library's kernels, forward and backward: the encoder through
`vsum_scorer_forward_train / vsum_scorer_backward`, `video_transform` through `vsum_linear_*` and the
three losses through `vsum_pretrain_losses_*` (packed layout; the repel term uses the O(N*d) form of
the reference's N x N cosine matrix, so no [bs,N,N] tensor exists).  PyTorch only carries the autograd
graph and owns the parameters.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import _cabi
from .simnet import SimNet, _al, pack_padded


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on packed rows (simnet_pretrain.py:80)."""

    @staticmethod
    def forward(ctx, x, w, b, mode):
        x, w, b = x.contiguous().float(), w.contiguous().float(), b.contiguous().float()
        M, K = x.shape
        N = w.shape[0]
        y = torch.empty((M, N), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            _cabi.check(_cabi.load().vsum_linear_forward(x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, N, K, mode,
                                                         _stream(x.device)), "vsum_linear_forward")
        ctx.save_for_backward(x, w)
        ctx.mode = mode
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        M, K = x.shape
        N = w.shape[0]
        dy = dy.contiguous().float()
        L = _cabi.load()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw, db = torch.empty_like(w), torch.empty(N, dtype=torch.float32, device=x.device)
        ws = torch.empty(L.vsum_linear_workspace_bytes(M, N, K) + 1024, dtype=torch.uint8, device=x.device)
        wp = _al(ws)
        with torch.cuda.device(x.device):
            _cabi.check(L.vsum_linear_backward(dy.data_ptr(), x.data_ptr(), w.data_ptr(), None if dx is None else dx.data_ptr(),
                                               dw.data_ptr(), db.data_ptr(), M, N, K, ctx.mode, wp, ws.numel() - (wp - ws.data_ptr()),
                                               _stream(x.device)), "vsum_linear_backward")
        return dx, dw, db, None


class _PretrainLossFn(torch.autograd.Function):
    """(scores [T,1], x512 [T,512]) -> float32[3] = (distillation, center, repel) (simnet_pretrain.py:82-100)."""

    @staticmethod
    def forward(ctx, scores, x512, cu, lens, n_pad, sharpening_t, video_rep, pen_entropy):
        dev = x512.device
        scores, x512 = scores.contiguous().float(), x512.contiguous().float()
        video_rep = video_rep.contiguous().float()
        T, B, max_len = x512.shape[0], len(lens), max(lens)
        L = _cabi.load()
        saved = torch.empty(L.vsum_pretrain_saved_bytes(T, B, max_len) + 1024, dtype=torch.uint8, device=dev)
        sp = _al(saved)
        losses = torch.empty(3, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _cabi.check(L.vsum_pretrain_losses_forward(scores.data_ptr(), x512.data_ptr(), cu.data_ptr(), B, T, max_len, n_pad,
                                                       float(sharpening_t), video_rep.data_ptr(), int(pen_entropy), losses.data_ptr(),
                                                       sp, saved.numel() - (sp - saved.data_ptr()), _stream(dev)),
                        "vsum_pretrain_losses_forward")
        ctx.save_for_backward(x512, cu)
        ctx.saved_blob, ctx.args = saved, (B, T, max_len, n_pad, float(sharpening_t), int(pen_entropy))
        return losses

    @staticmethod
    def backward(ctx, d_losses):
        x512, cu = ctx.saved_tensors
        B, T, max_len, n_pad, t, pen = ctx.args
        dev = x512.device
        d_losses = d_losses.contiguous().float()
        d_scores = torch.empty((T, 1), dtype=torch.float32, device=dev)
        d_x = torch.empty_like(x512)
        with torch.cuda.device(dev):
            _cabi.check(_cabi.load().vsum_pretrain_losses_backward(x512.data_ptr(), cu.data_ptr(), B, T, max_len, n_pad, t, pen,
                                                                   d_losses.data_ptr(), _al(ctx.saved_blob), d_scores.data_ptr(),
                                                                   d_x.data_ptr(), _stream(dev)), "vsum_pretrain_losses_backward")
        return d_scores, d_x, None, None, None, None, None, None


class PretrainModel(nn.Module):
    def __init__(self, feature_dim: int = 256, sparsity: float = 0.0, sharpening_t=0.4, **kwargs):
        super().__init__()
        self.feature_dim, self.sparsity, self.sharpening_t = feature_dim, sparsity, sharpening_t
        self.encoder = SimNet(sparsity=0., use_cls=False, d_model=feature_dim, **kwargs)
        self.video_transform = nn.Linear(feature_dim, 512)

    def forward(self, x, video_representation, mask=None, visualize_attention=None, pen_met="entropy"):
        if not x.is_cuda:
            raise _cabi.VsumError("vsum_b200 runs on CUDA devices only (no CPU fallback)")
        bs, n, _ = x.shape
        if isinstance(mask, Tensor):
            packed, cu, lens, _ = pack_padded(x, mask)
        else:
            packed, lens = x.reshape(bs * n, -1), [n] * bs
            cu = torch.arange(0, (bs + 1) * n, n, dtype=torch.int32, device=x.device)
        enc = self.encoder
        differentiable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if differentiable:
            scores, feats = enc.forward_packed_train(packed, cu, lens)          # model_score=True: feats = encoder output
        else:
            scores, feats = enc.forward_packed(packed.contiguous().float(), cu, lens)
        mode = 0 if enc.train_precision == "fp32" else 1
        x512 = _LinearFn.apply(feats, self.video_transform.weight, self.video_transform.bias, mode)
        losses = _PretrainLossFn.apply(scores, x512, cu, lens, n, self.sharpening_t, video_representation,
                                       pen_met == "entropy")
        return losses[0], losses[1], losses[2]
