"""`PretrainModel` with the reference's constructor / attributes / forward contract
(`src/model/simnet_pretrain.py:12-100`): `.encoder` is the CUDA `SimNet`, `.video_transform`
a `Linear(feature_dim, 512)`, `forward(x, video_representation, mask)` returns
`(loss, center_loss, repel_loss)`.

Round-1 status: the encoder (the hot part) runs on the CUDA kernels, forward AND backward (autograd
goes through `vsum_scorer_backward`, including the gradient w.r.t. the frame features it returns);
the three thin losses on top of it are PyTorch glue.  `repelling_loss` uses the O(N*d) algebraic
form of the reference's N x N cosine matrix, so no [bs,N,N] tensor exists.
Native loss kernels are section 8 row a10.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .simnet import SimNet


class PretrainModel(nn.Module):
    def __init__(self, feature_dim: int = 256, sparsity: float = 0.0, sharpening_t=0.4, **kwargs):
        super().__init__()
        self.feature_dim, self.sparsity, self.sharpening_t = feature_dim, sparsity, sharpening_t
        self.encoder = SimNet(sparsity=0., use_cls=False, d_model=feature_dim, **kwargs)
        self.video_transform = nn.Linear(feature_dim, 512)

    @staticmethod
    def cross_entropy_loss(x1, x2):
        return (-F.softmax(x2, dim=1) * torch.log(F.softmax(x1, dim=1))).mean()   # lines 35-41

    @staticmethod
    def entropy(x, mask=None):
        e = x * torch.log(x)
        if isinstance(mask, Tensor):
            e = e.masked_fill(mask, 0.)
        return e.mean(dim=1).mean()                                                # lines 43-47

    @staticmethod
    def repelling_loss(x: Tensor, mask):
        """mean_b[(1/N^2) * sum_{i != j} xh_i . xh_j] = (|sum_i xh_i|^2 - sum_i |xh_i|^2) / N^2
        with xh = x / (|x| + 1e-9), padded rows zeroed, N the padded length (lines 56-67)."""
        n = x.shape[1]
        if isinstance(mask, Tensor):
            x = x * (~mask).unsqueeze(2)
        xh = x / (x.norm(dim=2, keepdim=True) + 1e-9)
        total = xh.sum(dim=1).pow(2).sum(dim=1) - xh.pow(2).sum(dim=(1, 2))
        return (total / (n * n)).mean()

    def forward(self, x, video_representation, mask=None, visualize_attention=None, pen_met="entropy"):
        scores, frame_features = self.encoder(x, mask, model_score=True)
        frame_features = self.video_transform(frame_features)
        repel_loss = self.repelling_loss(frame_features, mask)
        mask3 = mask.unsqueeze(2)
        scores = scores.masked_fill(mask3, float("-inf"))
        mixture = F.softmax(scores / self.sharpening_t, dim=1)
        if pen_met == "entropy":
            center_loss = self.entropy(mixture + 1e-9, mask3)
        else:
            center_loss = torch.norm(mixture, dim=1).mean()
        pooled = torch.matmul(mixture.transpose(1, 2), frame_features).squeeze(1)
        loss = self.cross_entropy_loss(pooled, video_representation)
        return loss, center_loss, repel_loss
