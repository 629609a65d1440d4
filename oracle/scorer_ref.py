"""ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported by the product package.

Plain PyTorch fp32 CPU restatement of the reference frame scorer (`src/model/simnet.py`), written
functionally over a `state_dict` so it needs no module classes.  Eval mode only (every dropout is
the identity).  Kept as the floating-point reference for the CUDA scorer; tolerances live in the
tests (1e-5 for the fp32 kernels, 1e-2 for the bf16 tcgen05 kernels, as BASELINE.json states).

The arithmetic itself is third-party (PyTorch ATen / oneDNN, torch 2.11.0 in this image; the
reference pins no version).  Pinning: fixtures in `tests/golden/scorer_*.npz` are outputs of the
reference's own `SimNet` imported from `/root/reference/src` (`tests/golden/make_golden.py`).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def positional_table(n_rows: int, d_model: int) -> torch.Tensor:
    """`PositionalEncoding.__init__` (src/model/simnet.py:224-233): even cols sin, odd cols cos.
    The reference caps the table at 2000 rows (simnet.py:188); rows are independent of the cap,
    so a longer table is bit-identical on the first 2000 rows."""
    angle = torch.exp(-torch.arange(0, d_model, 2) * math.log(10000) / d_model)
    pos = torch.arange(0, n_rows).reshape(n_rows, 1)
    tab = torch.zeros((n_rows, d_model))
    tab[:, 0::2] = torch.sin(pos * angle)
    tab[:, 1::2] = torch.cos(pos * angle)
    return tab


def _lin(x, sd, prefix):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _scorer_forward(sd: dict, x: torch.Tensor, num_heads: int, key_padding_mask=None):
    """x [bs,N,1024] fp32 -> (logits [bs,N,C], feats [bs,N,d]).  `key_padding_mask` bool [bs,N],
    True = padded key (src/model/simnet.py:47-56,156-157)."""
    bs, n, _ = x.shape
    w_in = sd["embedding_layer.feature_transform.weight"]
    d = w_in.shape[0]
    h = _lin(x, sd, "embedding_layer.feature_transform")                  # simnet.py:211
    h = h + positional_table(n, d).unsqueeze(0)                            # simnet.py:236-238
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("encoder.module_list."))
    hd = d // num_heads
    scale = d ** -0.5                                                      # simnet.py:126 (d_model, not head_dim)
    for li in range(n_layers):
        p = f"encoder.module_list.{li}"
        q = _lin(h, sd, p + ".sa.q").view(bs, n, num_heads, hd).permute(0, 2, 1, 3)
        k = _lin(h, sd, p + ".sa.k").view(bs, n, num_heads, hd).permute(0, 2, 1, 3)
        v = _lin(h, sd, p + ".sa.v").view(bs, n, num_heads, hd).permute(0, 2, 1, 3)
        s = torch.matmul(q, k.transpose(2, 3)) * scale                     # simnet.py:155
        if key_padding_mask is not None:
            s = s.masked_fill(key_padding_mask.view(bs, 1, 1, n), float("-inf"))
        a = torch.matmul(F.softmax(s, dim=3), v)                           # simnet.py:158-160
        a = a.permute(0, 2, 1, 3).contiguous().view(bs, n, d)
        a = _lin(a, sd, p + ".sa.feature_projection")                      # simnet.py:163
        h = F.layer_norm(a + h, (d,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"])   # simnet.py:107
        m = _lin(F.relu(_lin(h, sd, p + ".mlp.fc1")), sd, p + ".mlp.fc2")  # simnet.py:181-182
        h = F.layer_norm(m + h, (d,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"])   # simnet.py:110
    return _lin(h, sd, "final_layer"), h                                   # simnet.py:42-45


scorer_forward = torch.no_grad()(_scorer_forward)
scorer_forward.__wrapped__ = _scorer_forward          # differentiable version for the gradient tests


def masked_mse(output, targets, mask):
    """`mse_with_mask_loss` (src/utils/utils.py:45-56): mean over bs*Nmax, pads contribute 0."""
    keep = (~mask).to(output.dtype)
    return (((output.squeeze(2) - targets) * keep) ** 2).mean()
