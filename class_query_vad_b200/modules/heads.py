"""DETR heads of the reference model (models/model.py:87-103,191-239) as a drop-in module for the TRAINING step:
`class_embed_b`, `bbox_embed` (shared with the decoder's box refinement, models/model.py:100-101) and Dropout(0.5) + channel mean
of the class tokens, through cqvad_heads_train_forward / _backward (fp32, like the reference, which disables autocast here).
Inference heads are part of cqvad_decoder_forward (DecoderEngine.forward(heads=True))."""
import ctypes

import torch
from torch import nn

from .. import _lib
from .decoder import MLP


class HeadsFunction(torch.autograd.Function):
    """apply(p_drop, seed, hs, cls_hs, refs, *8 parameters) -> (pred_logits [Lr,BT,nq,K], pred_boxes [Lr,BT,nq,4], pred_logits_b [Lr,BT,nq,3])"""

    @staticmethod
    def forward(ctx, p_drop, seed, hs, cls_hs, refs, *params):
        _lib.require_cuda(hs, cls_hs, refs)
        lib = _lib.lib()
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        hs_c, cls_c, refs_c = f32(hs), f32(cls_hs), f32(refs)
        W = [f32(p) for p in params]
        wtab = (ctypes.c_void_p * 8)(*[w.data_ptr() for w in W])
        Lr, BT, nq, K = cls_c.shape[:4]
        R = Lr * BT * nq
        dev = hs_c.device
        pl = torch.empty((Lr, BT, nq, K), dtype=torch.float32, device=dev)
        pb = torch.empty((Lr, BT, nq, 4), dtype=torch.float32, device=dev)
        plb = torch.empty((Lr, BT, nq, 3), dtype=torch.float32, device=dev)
        need = lib.cqvad_heads_train_workspace_bytes(R)
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
        p = _lib.ptr
        _lib.check(lib.cqvad_heads_train_forward(wtab, p(hs_c), p(cls_c), p(refs_c), R, K, float(p_drop), int(seed), p(pl), p(pb), p(plb),
                                                 p(ws), need, _lib.stream_ptr()))
        ctx.saved = (W, wtab, hs_c, refs_c, ws, need, R, K, float(p_drop), int(seed), cls_c.shape)
        ctx.dtypes = (hs.dtype, cls_hs.dtype, refs.dtype, [q.dtype for q in params])
        return pl, pb, plb

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_pl, g_pb, g_plb):
        W, wtab, hs_c, refs_c, ws, need, R, K, p_drop, seed, cls_shape = ctx.saved
        lib = _lib.lib()
        dev = hs_c.device
        f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()
        g_pl, g_pb, g_plb = f32(g_pl), f32(g_pb), f32(g_plb)
        g_hs = torch.empty_like(hs_c)
        g_refs = torch.empty_like(refs_c)
        g_cls = torch.empty(cls_shape, dtype=torch.float32, device=dev) if g_pl is not None else None
        gw = [torch.zeros_like(w) for w in W]
        gtab = (ctypes.c_void_p * 8)(*[g.data_ptr() for g in gw])
        p = _lib.ptr
        _lib.check(lib.cqvad_heads_train_backward(wtab, p(hs_c), p(refs_c), p(g_pl), p(g_pb), p(g_plb), R, K, p_drop, seed, p(g_hs),
                                                  p(g_cls), p(g_refs), gtab, p(ws), need, _lib.stream_ptr()))
        dh, dc, dr, dps = ctx.dtypes
        return (None, None, g_hs.to(dh), None if g_cls is None else g_cls.to(dc), g_refs.to(dr), *[g.to(d) for g, d in zip(gw, dps)])


class DETRHeads(nn.Module):
    """The head part of the reference `DETR` (models/model.py): same parameter names (`class_embed_b.*`, `bbox_embed.layers.*`), so
    the matching entries of a reference checkpoint load with strict=True.  forward(hs, cls_hs, reference) returns the reference's
    output dict {'pred_logits', 'pred_boxes', 'pred_logits_b', 'aux_outputs'} for the AVA single-frame configuration."""

    def __init__(self, hidden_dim=256, aux_loss=True, p_dropout=0.5):
        super().__init__()
        if hidden_dim != 256:
            raise ValueError("libcqvad heads: hidden_dim 256")
        self.class_embed_b = nn.Linear(hidden_dim, 3)
        self.bbox_embed = MLP(hidden_dim, hidden_dim, 4, 3)
        nn.init.constant_(self.bbox_embed.layers[-1].weight.data, 0)
        nn.init.constant_(self.bbox_embed.layers[-1].bias.data, 0)
        self.dropout = nn.Dropout(p_dropout)
        self.aux_loss = aux_loss
        self._calls = 0
        self.seed = 0x5EED

    def _params(self):
        L = self.bbox_embed.layers
        return [L[0].weight, L[0].bias, L[1].weight, L[1].bias, L[2].weight, L[2].bias, self.class_embed_b.weight, self.class_embed_b.bias]

    def forward(self, hs, cls_hs, reference):
        p = self.dropout.p if self.training else 0.0
        self._calls += 1          # a fresh Philox stream per call (the mask itself is never stored)
        pl, pb, plb = HeadsFunction.apply(p, (self.seed << 20) + self._calls, hs, cls_hs, reference, *self._params())
        out = {"pred_logits": pl[-1], "pred_boxes": pb[-1], "pred_logits_b": plb[-1]}
        if self.aux_loss:          # models/model.py:244-250
            out["aux_outputs"] = [{"pred_logits": a, "pred_boxes": b, "pred_logits_b": c} for a, b, c in zip(pl[:-1], pb[:-1], plb[:-1])]
        return out
