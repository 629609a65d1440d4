"""`SimNet` frame scorer with the reference's constructor, `forward` contract and `state_dict`
keys (reference `src/model/simnet.py:8-56`), executed by the sm_100a kernels behind the C ABI
(`vsum_scorer_forward`).  The nn.Module tree below only HOLDS parameters under the reference's
names so checkpoints load with `strict=True` in both directions; no PyTorch op computes the
scores.

Seeded-init parity: the reference's `Encoder` builds two extra blocks and throws them away
(`simnet.py:71-75`), which consumes init RNG before `final_layer` is drawn.  `_Encoder` replays
that consumption so `torch.manual_seed(s); SimNet(...)` yields bit-identical weights here and
there (tests/test_model_host.py).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Optional, Sequence

import torch
from torch import Tensor, nn

from .. import _cabi

_REF_TABLE_ROWS = 2000      # simnet.py:188 -- Embedding(max_len=2000) regardless of SimNet.max_len


def sinusoid_table(rows: int, d_model: int) -> Tensor:
    """fp32 table computed with the reference's exact expression order (simnet.py:224-231) on the
    host, so rows < 2000 are bit-identical to the reference buffer."""
    angle = torch.exp(-torch.arange(0, d_model, 2) * math.log(10000) / d_model)
    pos = torch.arange(0, rows).reshape(rows, 1)
    tab = torch.zeros((rows, d_model))
    tab[:, 0::2] = torch.sin(pos * angle)
    tab[:, 1::2] = torch.cos(pos * angle)
    return tab


class _PositionalEncoding(nn.Module):
    def __init__(self, d_model: int, rows: int):
        super().__init__()
        self.register_buffer("pos_embedding", sinusoid_table(rows, d_model).unsqueeze(0))


class _Embedding(nn.Module):
    def __init__(self, in_features: int, d_model: int, use_pos: bool, use_cls: bool):
        super().__init__()
        self.feature_transform = nn.Linear(in_features, d_model)
        if use_pos:
            self.positional_encoding = _PositionalEncoding(d_model, _REF_TABLE_ROWS)
        if use_cls:
            self.cls_token = nn.Parameter(torch.zeros((1, 1, d_model)))


class _Attention(nn.Module):
    def __init__(self, d_model: int):
        super().__init__()
        self.q = nn.Linear(d_model, d_model)
        self.k = nn.Linear(d_model, d_model)
        self.v = nn.Linear(d_model, d_model)
        self.feature_projection = nn.Linear(d_model, d_model)


class _FeedForward(nn.Module):
    def __init__(self, d_model: int, expand: int = 4):
        super().__init__()
        self.fc1 = nn.Linear(d_model, expand * d_model)
        self.fc2 = nn.Linear(expand * d_model, d_model)


class _Block(nn.Module):
    def __init__(self, d_model: int):
        super().__init__()
        self.sa = _Attention(d_model)
        self.mlp = _FeedForward(d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)


class _Encoder(nn.Module):
    def __init__(self, d_model: int, num_layers: int):
        super().__init__()
        self.module_list = nn.ModuleList([_Block(d_model) for _ in range(num_layers)])
        for _ in range(2):          # RNG replay of the two discarded blocks (simnet.py:72-74)
            _Block(d_model)
        self.module_score = nn.ModuleList([])


def _layer_tensors(blk):
    """Parameters of one encoder block in the field order of vsum_layer_weights / vsum_layer_grads."""
    return (blk.sa.q.weight, blk.sa.q.bias, blk.sa.k.weight, blk.sa.k.bias, blk.sa.v.weight, blk.sa.v.bias,
            blk.sa.feature_projection.weight, blk.sa.feature_projection.bias, blk.norm1.weight, blk.norm1.bias,
            blk.mlp.fc1.weight, blk.mlp.fc1.bias, blk.mlp.fc2.weight, blk.mlp.fc2.bias, blk.norm2.weight, blk.norm2.bias)


class _ScorerTrainFn(torch.autograd.Function):
    """Autograd bridge to vsum_scorer_forward_train / vsum_scorer_backward (arithmetic per `train_precision`)."""

    @staticmethod
    def forward(ctx, model, features, cu_seqlens, seqlens_host, drop_p, seed, *params):
        dev = features.device
        T, B = features.shape[0], len(seqlens_host)
        max_len = max(seqlens_host)
        L = _cabi.load()
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            model._sync_weights(max_len, dev, stream, train_only=True, force=True)   # every step: see _sync_weights
            h = model._handle
            _cabi.check(L.vsum_scorer_set_train_mode(h, model._train_mode()), "vsum_scorer_set_train_mode")
            tape = torch.empty(L.vsum_scorer_tape_bytes(h, T) + 1024, dtype=torch.uint8, device=dev)
            ws = model._workspace_for(L.vsum_scorer_train_workspace_bytes(h, T, B), dev)
            scores = torch.empty((T, model.num_classes), dtype=torch.float32, device=dev)
            feats = torch.empty((T, model.d_model), dtype=torch.float32, device=dev)
            tp, wp = _al(tape), _al(ws)
            _cabi.check(L.vsum_scorer_forward_train(h, features.data_ptr(), cu_seqlens.data_ptr(), B, T, max_len,
                                                    float(drop_p), int(seed), scores.data_ptr(), feats.data_ptr(),
                                                    tp, tape.numel() - (tp - tape.data_ptr()), wp,
                                                    ws.numel() - (wp - ws.data_ptr()), stream), "vsum_scorer_forward_train")
        ctx.model, ctx.tape, ctx.args = model, tape, (features, cu_seqlens, B, T, max_len, float(drop_p), int(seed))
        ctx.set_materialize_grads(False)
        return scores, feats

    @staticmethod
    def backward(ctx, d_scores, d_feats):
        model, tape = ctx.model, ctx.tape
        features, cu_seqlens, B, T, max_len, drop_p, seed = ctx.args
        dev = features.device
        L = _cabi.load()
        # one flat, zeroed gradient buffer; every .grad is a view of it (sharding.allreduce_gradients reduces it
        # in place) and q|k|v of a layer are adjacent so the fused QKV weight gradient is written in place
        lay = model._grad_layout()
        dp = getattr(model, "_dp", None)
        n_grads = lay["n_grads"]
        flat = torch.zeros(n_grads + (dp.ext_n if dp is not None else 0), dtype=torch.float32, device=dev)   # + sharding.DataParallel's extras
        grads = [flat[o:o + n].view(shape) for o, n, shape in lay["views"]]
        if d_scores is None:
            d_scores = torch.zeros((T, model.num_classes), dtype=torch.float32, device=dev)
        d_scores = d_scores.contiguous().float()
        d_feats = None if d_feats is None else d_feats.contiguous().float()
        g = _cabi.ScorerGrads()
        base = flat.data_ptr()
        ptrs = [base + 4 * o for o, _, _ in lay["views"]]           # same order as SimNet._train_params
        g.embed_w, g.embed_b = ptrs[0], ptrs[1]
        k = 2
        for i in range(model.num_layers):
            gl = g.layers[i]
            for name in _cabi._LAYER_FIELDS:
                setattr(gl, name, ptrs[k])
                k += 1
        g.final_w, g.final_b = ptrs[k], ptrs[k + 1]
        g.pre_zeroed = 1
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            # (the handle holds the weights of this step's forward: nothing to refresh between forward and backward)
            _cabi.check(L.vsum_scorer_set_train_mode(model._handle, model._train_mode()), "vsum_scorer_set_train_mode")
            ws = model._workspace_for(L.vsum_scorer_train_workspace_bytes(model._handle, T, B), dev)
            wp = _al(ws)
            args = (model._handle, features.data_ptr(), cu_seqlens.data_ptr(), B, T, max_len,
                    drop_p, seed, d_scores.data_ptr(), None if d_feats is None else d_feats.data_ptr(),
                    _al(tape), C.byref(g), wp, ws.numel() - (wp - ws.data_ptr()), stream)
            if dp is None:
                _cabi.check(L.vsum_scorer_backward(*args), "vsum_scorer_backward")
            else:
                # data-parallel step (sharding.DataParallel): the flat buffer is all-reduced bucket by bucket on the communication
                # stream while the backward of the earlier layers is still running
                dp._begin(flat, n_grads, lay["embed_n"], lay["layer_n"], model.num_layers, T)
                hook = _cabi.GRAD_BUCKET_HOOK(lambda _user, bucket: dp._bucket_ready(int(bucket)))
                _cabi.check(L.vsum_scorer_backward_hooked(*args, hook, None), "vsum_scorer_backward_hooked")
        return (None, None, None, None, None, None, *grads)


def pack_padded(x: Tensor, mask: Tensor):
    """Padded batch [bs,n,F] + bool mask [bs,n] (True = padded frame, src/train.py:118) -> packed rows [T,F],
    cu_seqlens int32[bs+1] on the device, host lengths, and the keep mask."""
    bs, n = mask.shape
    dev = x.device
    keep = ~mask
    lens = keep.sum(dim=1)
    ramp = torch.arange(n, device=dev).unsqueeze(0) < lens.unsqueeze(1)
    lens_host = [int(v) for v in lens.tolist()]
    if not bool((ramp == keep).all()):
        raise _cabi.VsumError("only suffix padding (pad_sequence layout, dataset.py:157-161) is supported")
    packed = x[keep]                                               # [T,F] gather of the valid frames
    cu = torch.zeros(bs + 1, dtype=torch.int32, device=dev)
    cu[1:] = lens.cumsum(0).to(torch.int32)
    return packed, cu, lens_host, keep


def _al(t: Tensor) -> int:
    return (t.data_ptr() + 1023) // 1024 * 1024


class SimNet(nn.Module):
    """Drop-in for the reference `SimNet`.  Extra, optional attribute: `precision` ("bf16" uses
    the tcgen05 kernels and needs d_model=256, heads=4; "fp32" uses the fp32 SIMT kernels)."""

    def __init__(self, num_heads: int = 8, d_model: int = 512, num_layers: int = 4,
                 sparsity: float = 0.5, use_cls: bool = False, dropout: float = 0.2,
                 num_classes: int = 1, use_pos: bool = True, max_len=2500):
        super().__init__()
        if use_cls:
            raise NotImplementedError(
                "use_cls=True is never used by the reference's callers (train.py:33, "
                "simnet_pretrain.py:30) and hard-codes a CUDA tensor there (simnet.py:49); not built")
        self.num_heads, self.d_model, self.num_layers = num_heads, d_model, num_layers
        self.sparsity, self.use_cls, self.max_len = sparsity, use_cls, max_len
        self.num_classes, self.in_features, self.use_pos = num_classes, 1024, use_pos
        self.dropout = dropout
        self.embedding_layer = _Embedding(self.in_features, d_model, use_pos, use_cls)
        self.encoder = _Encoder(d_model, num_layers)
        self.final_layer = nn.Linear(d_model, num_classes)

        tc05 = d_model == 256 and num_heads == 4 and num_classes == 1
        self.precision = os.environ.get("VSUM_PRECISION", "bf16" if tc05 else "fp32")
        # TRAINING path: "bf16" = every contraction on tcgen05 (tf32 linears, bf16 wgrad and attention -- the
        # counterpart of the reference's autocast, train.py:120), "tf32" = tensor-core linears with fp32
        # attention, "fp32" = SIMT kernels at reference accuracy
        self.train_precision = os.environ.get("VSUM_TRAIN_PRECISION", "bf16" if tc05 else "fp32")
        self._handle = None
        self._dp = None                      # sharding.DataParallel attaches itself here
        self._weights_key = None
        self._weights_key_train = None
        self._table: Optional[Tensor] = None
        self._workspace: Optional[Tensor] = None

    # ------------------------------------------------------------------ C-ABI plumbing
    def _mode(self) -> int:
        if self.precision not in ("bf16", "fp32"):
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {self.precision!r}")
        return _cabi.MODE_BF16 if self.precision == "bf16" else _cabi.MODE_FP32

    def _ensure_handle(self):
        if self._handle is None:
            cfg = _cabi.ScorerConfig(self.d_model, self.num_heads, self.num_layers, 4 * self.d_model,
                                     self.in_features, self.num_classes, int(self.use_pos), 0)
            h = C.c_void_p()
            _cabi.check(_cabi.load().vsum_scorer_create(C.byref(h), C.byref(cfg)), "vsum_scorer_create")
            self._handle = h
        return self._handle

    def __del__(self):
        try:
            if getattr(self, "_handle", None) is not None:
                _cabi.load().vsum_scorer_destroy(self._handle)
        except Exception:
            pass

    def _table_for(self, rows_needed: int, device) -> Tensor:
        """Positional rows.  The registered buffer keeps the reference shape [1,2000,d]; longer
        videos (BASELINE configs go to N=8192, where the reference itself raises) extend the same
        formula."""
        buf = self.embedding_layer.positional_encoding.pos_embedding
        if rows_needed <= buf.shape[1]:
            return buf[0]
        if self._table is None or self._table.shape[0] < rows_needed or self._table.device != device:
            rows = 1 << (rows_needed - 1).bit_length()
            self._table = sinusoid_table(rows, self.d_model).to(device)
            self._weights_key = self._weights_key_train = None
        return self._table

    def _param_list(self):
        """Parameters in module order, cached (the module tree is fixed after construction; walking it costs ~0.2 ms)."""
        pl = self.__dict__.get("_params_cache")
        if pl is None:
            pl = list(self.parameters())
            self.__dict__["_params_cache"] = pl
        return pl

    def _apply(self, fn, *args, **kwargs):
        """`.to()` / `.cuda()` / `.float()` may replace parameter tensors: drop everything cached about them."""
        for k in ("_params_cache", "_train_params_cache", "_grad_layout_cache", "_weights_struct"):
            self.__dict__.pop(k, None)
        self._weights_key = self._weights_key_train = None
        return super()._apply(fn, *args, **kwargs)

    def _grad_layout(self):
        """Offsets of every parameter's gradient inside the flat buffer `_ScorerTrainFn.backward` hands out, in
        `_train_params` order: [embedding | per layer q_w k_w v_w q_b k_b v_b, then the rest | head]."""
        lay = self.__dict__.get("_grad_layout_cache")
        if lay is not None:
            return lay
        params = list(self._train_params())
        names = ["embed_w", "embed_b"] + [f"{i}.{f}" for i in range(self.num_layers) for f in _cabi._LAYER_FIELDS] + ["final_w", "final_b"]
        sizes = dict(zip(names, (p.numel() for p in params)))
        order = ["embed_w", "embed_b"]
        for i in range(self.num_layers):
            first = [f"{i}.{f}" for f in ("q_w", "k_w", "v_w", "q_b", "k_b", "v_b")]
            order += first + [f"{i}.{f}" for f in _cabi._LAYER_FIELDS if f"{i}.{f}" not in first]
        order += ["final_w", "final_b"]
        offs, off = {}, 0
        for k in order:
            offs[k] = off
            off += sizes[k]
        lay = dict(n_grads=off, views=[(offs[k], sizes[k], tuple(p.shape)) for k, p in zip(names, params)],
                   embed_n=sizes["embed_w"] + sizes["embed_b"], layer_n=sum(sizes[f"0.{f}"] for f in _cabi._LAYER_FIELDS))
        self.__dict__["_grad_layout_cache"] = lay
        return lay

    def mark_weights_dirty(self):
        """Force the next forward to hand the parameters to the kernels again.  Inference calls detect updates through the
        tensors' version counters; an update that does not bump them (a fused optimizer, a CUDA-graph replay, writes through
        `.data`) outside the differentiable forward -- which always refreshes -- has to be announced here."""
        self._weights_key = self._weights_key_train = None

    def _sync_weights(self, max_len: int, device, stream: int, train_only: bool = False, force: bool = False):
        """Hand the current parameter values to the handle.  `train_only`: the per-step refresh of a training loop -- fp32
        copies and transposes only (VSUM_WEIGHTS_TRAIN_ONLY); the bf16 inference copies are rebuilt by the next inference
        call.  `force`: do not trust the version counters (the differentiable forward: `torch.optim.Adam(fused=True)` updates
        the parameters without bumping them, and a stale copy would silently train on the old weights)."""
        params = self._param_list()
        table = self._table_for(max_len, device) if self.use_pos else None
        key = (tuple((p.data_ptr(), p._version) for p in params),
               None if table is None else (table.data_ptr(), table.shape[0]))
        if not force and (key == self._weights_key or (train_only and key == self._weights_key_train)):
            return
        ptr_key = (tuple(k[0] for k in key[0]), key[1], str(device))
        cached = self.__dict__.get("_weights_struct")
        if cached is not None and cached[0] == ptr_key:      # same tensors as last time: the filled struct is still right
            _cabi.check(_cabi.load().vsum_scorer_load_weights_ex(self._ensure_handle(), C.byref(cached[1]),
                                                                 _cabi.WEIGHTS_TRAIN_ONLY if train_only else 0, C.c_void_p(stream)),
                        "vsum_scorer_load_weights_ex")
            self._weights_key_train = key
            self._weights_key = None if train_only else key
            return
        for p in params:
            if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
                raise _cabi.VsumError("SimNet parameters must be contiguous fp32 tensors on the input's device")
        w = _cabi.ScorerWeights()
        emb = self.embedding_layer.feature_transform
        w.embed_w, w.embed_b = emb.weight.data_ptr(), emb.bias.data_ptr()
        if table is not None:
            table = table.contiguous()
            w.pos_table, w.pos_rows = table.data_ptr(), table.shape[0]
        w.final_w, w.final_b = self.final_layer.weight.data_ptr(), self.final_layer.bias.data_ptr()
        for i, blk in enumerate(self.encoder.module_list):
            lw = w.layers[i]
            lw.q_w, lw.q_b = blk.sa.q.weight.data_ptr(), blk.sa.q.bias.data_ptr()
            lw.k_w, lw.k_b = blk.sa.k.weight.data_ptr(), blk.sa.k.bias.data_ptr()
            lw.v_w, lw.v_b = blk.sa.v.weight.data_ptr(), blk.sa.v.bias.data_ptr()
            lw.o_w, lw.o_b = blk.sa.feature_projection.weight.data_ptr(), blk.sa.feature_projection.bias.data_ptr()
            lw.ln1_g, lw.ln1_b = blk.norm1.weight.data_ptr(), blk.norm1.bias.data_ptr()
            lw.fc1_w, lw.fc1_b = blk.mlp.fc1.weight.data_ptr(), blk.mlp.fc1.bias.data_ptr()
            lw.fc2_w, lw.fc2_b = blk.mlp.fc2.weight.data_ptr(), blk.mlp.fc2.bias.data_ptr()
            lw.ln2_g, lw.ln2_b = blk.norm2.weight.data_ptr(), blk.norm2.bias.data_ptr()
        _cabi.check(_cabi.load().vsum_scorer_load_weights_ex(self._ensure_handle(), C.byref(w),
                                                             _cabi.WEIGHTS_TRAIN_ONLY if train_only else 0, C.c_void_p(stream)),
                    "vsum_scorer_load_weights_ex")
        self.__dict__["_weights_struct"] = (ptr_key, w, table)          # (the table tensor is kept alive with the struct that points at it)
        self._weights_key_train = key
        self._weights_key = None if train_only else key

    def _workspace_for(self, nbytes: int, device) -> Tensor:
        ws = self._workspace
        if ws is None or ws.numel() < nbytes + 1024 or ws.device != device:
            ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
            self._workspace = ws
        return ws

    def _train_mode(self) -> int:
        if self.train_precision not in ("bf16", "tf32", "fp32"):
            raise ValueError(f"train_precision must be 'bf16', 'tf32' or 'fp32', got {self.train_precision!r}")
        return {"fp32": 0, "tf32": 1, "bf16": 2}[self.train_precision]

    def _train_params(self):
        """Parameters in the order _ScorerTrainFn returns their gradients."""
        tp = self.__dict__.get("_train_params_cache")
        if tp is not None:
            return iter(tp)
        tp = list(self._train_params_walk())
        self.__dict__["_train_params_cache"] = tp
        return iter(tp)

    def _train_params_walk(self):
        emb = self.embedding_layer.feature_transform
        yield emb.weight
        yield emb.bias
        for blk in self.encoder.module_list:
            yield from _layer_tensors(blk)
        yield self.final_layer.weight
        yield self.final_layer.bias

    def forward_packed_train(self, features: Tensor, cu_seqlens: Tensor, seqlens_host: Sequence[int]):
        """Differentiable packed forward (native forward with tape + native backward).  Dropout follows
        `self.training` / `self.dropout` like nn.Dropout in the reference (simnet.py:107,110,159,181)."""
        if not features.is_cuda:
            raise _cabi.VsumError("vsum_b200 runs on CUDA devices only (no CPU fallback)")
        if self.training and float(getattr(self, "sparsity", 0.0) or 0.0) > 0.0:
            # the reference drops (embedding + positional encoding) with p = sparsity (simnet.py:204-206, 235-237); every caller
            # of the reference passes 0 (train.py:32, simnet_pretrain.py:30) and this path has no dropout site there
            raise _cabi.VsumError("SimNet(sparsity > 0) in training mode is not built (the reference's callers use sparsity=0); "
                                  "construct the model with sparsity=0. or call .eval()")
        drop_p = float(self.dropout) if self.training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        return _ScorerTrainFn.apply(self, features.contiguous().float(), cu_seqlens, list(seqlens_host), drop_p, seed,
                                    *self._train_params())

    # ------------------------------------------------------------------ packed entry point
    @torch.no_grad()
    def forward_packed(self, features: Tensor, cu_seqlens: Tensor, seqlens_host: Sequence[int],
                       apply_sigmoid: bool = False, want_feats: bool = True,
                       scores_out: Optional[Tensor] = None):
        """features [T,1024] fp32 on a CUDA device, rows of video v = [cu[v], cu[v+1]).
        Returns (scores [T,num_classes] fp32, feats [T,d_model] fp32 | None).  `scores_out` lets a
        pipelined caller supply the (contiguous fp32 [T,num_classes]) output buffer.
        bf16 features (a pack written with `features_bf16`) are accepted by the bf16 scorer: the feature GEMM
        then runs in bf16 like the rest of the network instead of tf32."""
        if not features.is_cuda:
            raise _cabi.VsumError("vsum_b200 runs on CUDA devices only (no CPU fallback); move the input with .cuda()")
        bf16_in = features.dtype == torch.bfloat16
        if bf16_in and self.precision != "bf16":
            raise ValueError("bf16 features need the bf16 scorer (precision='bf16'); the fp32 mode reads float32 features")
        if features.dtype not in (torch.float32, torch.bfloat16) or features.dim() != 2 or features.shape[1] != self.in_features:
            raise ValueError(f"features must be float32 (or bfloat16) [T,{self.in_features}], got {tuple(features.shape)} {features.dtype}")
        features = features.contiguous()
        dev = features.device
        T, B = features.shape[0], len(seqlens_host)
        max_len = max(seqlens_host) if B else 0
        if scores_out is not None:
            if scores_out.shape != (T, self.num_classes) or scores_out.dtype != torch.float32 or not scores_out.is_contiguous():
                raise ValueError("scores_out must be a contiguous float32 [T,num_classes] tensor")
            scores = scores_out
        else:
            scores = torch.empty((T, self.num_classes), dtype=torch.float32, device=dev)
        feats = torch.empty((T, self.d_model), dtype=torch.float32, device=dev) if want_feats else None
        if T == 0:
            return scores, feats
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            self._sync_weights(max_len, dev, stream)
            L, mode = _cabi.load(), (_cabi.MODE_BF16_FEATURES if bf16_in else self._mode())
            need = L.vsum_scorer_workspace_bytes(self._handle, T, B, mode)
            ws = self._workspace_for(need, dev)
            ws_ptr = (ws.data_ptr() + 1023) // 1024 * 1024
            _cabi.check(L.vsum_scorer_forward(
                self._handle, features.data_ptr(), cu_seqlens.data_ptr(), B, T, max_len, mode,
                int(apply_sigmoid), scores.data_ptr(), feats.data_ptr() if want_feats else None,
                ws_ptr, ws.numel() - (ws_ptr - ws.data_ptr()), stream), "vsum_scorer_forward")
        return scores, feats

    # ------------------------------------------------------------------ reference contract
    def forward(self, x, mask=None, vis_attention=None, model_score=False):
        """x [bs,n,1024]; mask bool [bs,n], True = padded frame (src/train.py:118).  A non-tensor
        mask is ignored like the reference does (simnet.py:38, train.py:162).  Returns
        (scores [bs,n,num_classes], feats [bs,n,d_model]); with `model_score=True` the second
        element is the same tensor because the reference's score stack is empty (simnet.py:80-83)."""
        bs, n, _ = x.shape
        dev = x.device
        # autograd on (train_step, train.py:111-131): fp32 kernels with the native backward
        differentiable = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        run = self.forward_packed_train if differentiable else self.forward_packed
        if isinstance(mask, Tensor):
            packed, cu, lens_host, keep = pack_padded(x, mask)
            s, f = run(packed, cu, lens_host)
            idx = keep.reshape(-1).nonzero().squeeze(1)            # padded rows stay 0 (the loss masks them)
            scores = torch.zeros((bs * n, self.num_classes), dtype=torch.float32, device=dev).index_copy(0, idx, s)
            feats = torch.zeros((bs * n, self.d_model), dtype=torch.float32, device=dev).index_copy(0, idx, f)
            return scores.view(bs, n, self.num_classes), feats.view(bs, n, self.d_model)
        cu = torch.arange(0, (bs + 1) * n, n, dtype=torch.int32, device=dev)
        s, f = run(x.reshape(bs * n, -1), cu, [n] * bs)
        return s.view(bs, n, self.num_classes), f.view(bs, n, self.d_model)
