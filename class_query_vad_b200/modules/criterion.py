"""Drop-in mirrors of the reference's AVA matcher / criterion / post-processor on top of
cqvad_criterion_ava / cqvad_postprocess_ava: the assignment, the four losses, the weighted total and its gradient are computed on
the device without a host round trip (the reference copies the cost matrix to the host and calls scipy per decoder output,
models/detr/matcher.py:72-77).

  HungarianMatcherAVA   models/detr/matcher.py:13-78        same constructor, forward(outputs, targets) -> [(index_i, index_j)]
  SetCriterionAVA       models/detr/criterion.py:17-224     same constructor, forward(outputs, targets) -> dict of losses
  PostProcessAVA        models/detr/criterion.py:738-773    forward(outputs, target_sizes) -> (scores, boxes, person) numpy arrays

Targets keep the reference format: a list with one dict per clip, "boxes" [n, 5] (column 0 = key-frame id, dropped as in
matcher.py:60) and "labels" [n, K] multi-hot.  `pack_targets` pads them once into the dense arrays of include/cqvad.h."""
import ctypes

import torch
from torch import nn

from .. import _lib

_LOSS_KEYS = ("loss_ce", "loss_bbox", "loss_giou", "loss_ce_b")


def pack_targets(targets, K, device):
    """list of {"boxes": [n,5], "labels": [n,K]} -> (tgt_boxes [B,maxT,4], tgt_labels [B,maxT,K], n_tgt [B] int32) on `device`."""
    B = len(targets)
    n = [int(t["boxes"].shape[0]) for t in targets]
    maxT = max(max(n), 1)
    tb = torch.zeros((B, maxT, 4), dtype=torch.float32)
    tl = torch.zeros((B, maxT, K), dtype=torch.float32)
    for b, t in enumerate(targets):
        if n[b]:
            tb[b, :n[b]] = t["boxes"].detach().float().cpu()[:, 1:]
            tl[b, :n[b]] = t["labels"].detach().float().cpu()
    return tb.to(device), tl.to(device), torch.tensor(n, dtype=torch.int32).to(device)


def _cfg(cost_class=1.0, cost_bbox=1.0, cost_giou=1.0, weight_dict=None, pos_weight=1.0, eos_coef=0.1, alpha=0.25, gamma=2.0,
         smoothing=0.1):
    wd = weight_dict or {}
    return _lib.CriterionCfg(float(cost_class), float(cost_bbox), float(cost_giou), float(wd.get("loss_ce", 0.0)),
                             float(wd.get("loss_bbox", 0.0)), float(wd.get("loss_giou", 0.0)), float(wd.get("loss_ce_b", 0.0)),
                             float(pos_weight), float(eos_coef), float(alpha), float(gamma), float(smoothing or 0.0))


def criterion_ava(cfg, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt, want_grad=True):
    """One launch sequence of cqvad_criterion_ava.  Returns (losses [16] fp32 device tensor, match [B,nq] int32, grads or None)."""
    _lib.require_cuda(pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt)
    pl, pb, plb = (t.detach().float().contiguous() for t in (pred_logits, pred_boxes, pred_logits_b))
    B, nq, K = pl.shape
    maxT = tgt_boxes.shape[1]
    dev = pl.device
    match = torch.empty((B, nq), dtype=torch.int32, device=dev)
    losses = torch.zeros(16, dtype=torch.float32, device=dev)
    grads = (torch.empty_like(pl), torch.empty_like(pb), torch.empty_like(plb)) if want_grad else (None, None, None)
    L = _lib.lib()
    nbytes = L.cqvad_criterion_ava_workspace_bytes(B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    p = _lib.ptr
    _lib.check(L.cqvad_criterion_ava(ctypes.byref(cfg), p(pl), p(pb), p(plb), p(tgt_boxes), p(tgt_labels), p(n_tgt), B, nq, K, maxT,
                                     p(match), p(losses), p(grads[0]), p(grads[1]), p(grads[2]), p(ws), nbytes, _lib.stream_ptr()))
    return losses, match, (grads if want_grad else None)


class _CriterionFunction(torch.autograd.Function):
    """losses[0..5] as a differentiable function of the three prediction tensors; the backward scales the stored gradient of the
    weighted total -- valid for any linear combination of the four losses taken with the weight_dict coefficients, which is the
    only use the reference makes of them (train.py:148).  Each loss_k is returned as a separate output; d/d loss_k routes the
    per-loss share, so a caller who re-weights the dict still gets the right gradient."""

    @staticmethod
    def forward(ctx, cfgs, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt):
        per_loss = []
        base, singles = cfgs
        losses, match, _ = criterion_ava(base, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt, want_grad=False)
        for c in singles:          # gradient of each loss alone (unit weight): four more tiny launches of the gradient kernel
            _, _, g = criterion_ava(c, pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt, want_grad=True)
            per_loss.append(g)
        ctx.per_loss = per_loss
        ctx.dtypes = (pred_logits.dtype, pred_boxes.dtype, pred_logits_b.dtype)
        outs = tuple(losses[i].clone() for i in range(4)) + (losses[5].clone(), match)
        ctx.mark_non_differentiable(outs[4], outs[5])
        return outs

    @staticmethod
    def backward(ctx, g_ce, g_bbox, g_giou, g_ce_b, _g_err, _g_match):
        gs = (g_ce, g_bbox, g_giou, g_ce_b)
        out = [None, None, None]
        for k, g in enumerate(gs):
            if g is None:
                continue
            for i in range(3):
                term = ctx.per_loss[k][i] * g
                out[i] = term if out[i] is None else out[i] + term
        out = [None if o is None else o.to(dt) for o, dt in zip(out, ctx.dtypes)]
        return (None, out[0], out[1], out[2], None, None, None)


class HungarianMatcherAVA(nn.Module):
    def __init__(self, cost_class: float = 1, cost_bbox: float = 1, cost_giou: float = 1, binary_loss: bool = False,
                 before: bool = False, clip_len: int = 32):
        super().__init__()
        self.cost_class, self.cost_bbox, self.cost_giou = cost_class, cost_bbox, cost_giou
        self.binary_loss, self.before, self.clip_len = binary_loss, before, clip_len
        assert cost_class != 0 or cost_bbox != 0 or cost_giou != 0, "all costs cant be 0"

    @torch.no_grad()
    def match_dense(self, outputs, tgt_boxes, tgt_labels, n_tgt):
        cfg = _cfg(self.cost_class, self.cost_bbox, self.cost_giou)
        _, match, _ = criterion_ava(cfg, outputs["pred_logits"], outputs["pred_boxes"], outputs["pred_logits_b"], tgt_boxes,
                                    tgt_labels, n_tgt, want_grad=False)
        return match

    @torch.no_grad()
    def forward(self, outputs, targets):
        """-> [(index_i, index_j)] per clip as the reference (int64, ordered by query index like scipy's row_ind)."""
        K = outputs["pred_logits"].shape[-1]
        tb, tl, nt = pack_targets(targets, K, outputs["pred_logits"].device)
        m = self.match_dense(outputs, tb, tl, nt).cpu()
        res = []
        for b in range(m.shape[0]):
            qi = torch.nonzero(m[b] >= 0, as_tuple=False).flatten()
            res.append((qi.to(torch.int64), m[b, qi].to(torch.int64)))
        return res


class SetCriterionAVA(nn.Module):
    def __init__(self, weight, num_classes, num_queries, matcher, weight_dict, eos_coef, losses, data_file, evaluation=False,
                 label_smoothing_alpha=0.1):
        super().__init__()
        if evaluation:
            raise NotImplementedError("SetCriterionAVA(evaluation=True) (plain BCE, criterion.py:92-93) is not built")
        unknown = set(losses) - {"labels", "boxes"}
        if unknown:
            raise NotImplementedError(f"losses {sorted(unknown)} are not built (the shipped yamls use ['labels', 'boxes'])")
        self.weight, self.num_classes, self.num_queries = weight, num_classes, num_queries
        self.matcher, self.weight_dict, self.eos_coef, self.losses, self.data_file = matcher, weight_dict, eos_coef, losses, data_file
        self.register_buffer("empty_weight", torch.tensor([1.0, 1.0, float(eos_coef)]))
        self.focal_loss_alpha, self.focal_loss_gamma = 0.25, 2.0
        self.label_smoothing_alpha = 0.1          # the reference ignores the constructor argument (criterion.py:48)

    def _cfgs(self):
        m = self.matcher
        kw = dict(cost_class=m.cost_class, cost_bbox=m.cost_bbox, cost_giou=m.cost_giou, pos_weight=self.weight,
                  eos_coef=self.eos_coef, alpha=self.focal_loss_alpha, gamma=self.focal_loss_gamma,
                  smoothing=self.label_smoothing_alpha)
        base = _cfg(weight_dict=self.weight_dict, **kw)
        singles = [_cfg(weight_dict={k: 1.0}, **kw) for k in _LOSS_KEYS]
        return base, singles

    def forward_dense(self, outputs, tgt_boxes, tgt_labels, n_tgt):
        """The losses of one decoder output on pre-packed targets.  Returns (dict, match [B,nq] int32)."""
        ce, bbox, giou, ce_b, err, match = _CriterionFunction.apply(self._cfgs(), outputs["pred_logits"], outputs["pred_boxes"],
                                                                    outputs["pred_logits_b"], tgt_boxes, tgt_labels, n_tgt)
        return {"loss_ce": ce, "loss_ce_b": ce_b, "class_error": err, "loss_bbox": bbox, "loss_giou": giou}, match

    def forward(self, outputs, targets):
        K = outputs["pred_logits"].shape[-1]
        tb, tl, nt = pack_targets(targets, K, outputs["pred_logits"].device)
        losses, _ = self.forward_dense({k: v for k, v in outputs.items() if k != "aux_outputs"}, tb, tl, nt)
        for i, aux in enumerate(outputs.get("aux_outputs", [])):            # criterion.py:209-223 (logging only: not in weight_dict)
            l_dict, _ = self.forward_dense(aux, tb, tl, nt)
            l_dict.pop("class_error")
            losses.update({f"{k}_{i}": v for k, v in l_dict.items()})
        return losses

    def total_and_grads(self, outputs, tgt_boxes, tgt_labels, n_tgt):
        """Training fast path: the weighted total of train.py:148 and its gradient in ONE launch sequence, no autograd graph.
        Returns (losses [16] device tensor -- [4] is the total --, match, (g_logits, g_boxes, g_logits_b))."""
        return criterion_ava(self._cfgs()[0], outputs["pred_logits"], outputs["pred_boxes"], outputs["pred_logits_b"], tgt_boxes,
                             tgt_labels, n_tgt, want_grad=True)


class PostProcessAVA(nn.Module):
    _entry = "cqvad_postprocess_ava"

    @torch.no_grad()
    def detections(self, outputs, target_sizes):
        """[B, nq, K+5] device tensor = [sigmoid scores | xyxy boxes in pixels | person probability] (what dist.gather_detections ships)."""
        pl = outputs["pred_logits"].detach().float().contiguous()
        pb = outputs["pred_boxes"].detach().float().contiguous()
        plb = outputs["pred_logits_b"].detach().float().contiguous()
        _lib.require_cuda(pl, pb, plb)
        assert len(pl) == len(target_sizes) and target_sizes.shape[1] == 2
        ts = target_sizes.detach().to(device=pl.device, dtype=torch.float32).contiguous()
        B, nq, K = pl.shape
        det = torch.empty((B, nq, K + 5), dtype=torch.float32, device=pl.device)
        p = _lib.ptr
        _lib.check(getattr(_lib.lib(), self._entry)(p(pl), p(pb), p(plb), p(ts), p(det), B, nq, K, _lib.stream_ptr()))
        return det

    @torch.no_grad()
    def forward(self, outputs, target_sizes):
        det = self.detections(outputs, target_sizes).cpu().numpy()
        K = outputs["pred_logits"].shape[-1]
        return det[..., :K], det[..., K:K + 4], det[..., K + 4:K + 5]


class PostProcessUCF(PostProcessAVA):
    """models/detr/criterion.py:775-809: class scores gated by the person probability (outputs flattened to frames: [B*T', nq, .])."""
    _entry = "cqvad_postprocess_ucf"


class PostProcessJHMDB(PostProcessUCF):
    """models/detr/criterion.py:811-846 (identical arithmetic to PostProcessUCF)."""
