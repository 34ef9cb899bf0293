"""TEST INFRASTRUCTURE ONLY -- numpy (fp64) restatement of the AVA loss / matcher / post-process of the reference
(SURVEY.md section 8f row 4).  Pinned against the reference's own modules by tests/golden/criterion_*.npz
(oracle/make_golden_criterion.py); nothing in the product imports this file.

  generalized_box_iou / box_cxcywh_to_xyxy   utils/box_ops.py:9-14,40-108
  match_ava                                  models/detr/matcher.py:39-78   (HungarianMatcherAVA.forward)
  criterion_ava                              models/detr/criterion.py:50-105,119-138,184-224 (SetCriterionAVA) and the
                                             weighted total of train.py:148
  sigmoid focal loss                         models/detr/segmentation.py:200-229
  postprocess_ava                            models/detr/criterion.py:740-773
"""
import numpy as np
from scipy.optimize import linear_sum_assignment

DEFAULT_CFG = dict(cost_class=12.0, cost_bbox=5.0, cost_giou=2.0,          # configuration/AVA22_ViT-B.yaml:71-86
                   w_ce=10.0, w_bbox=5.0, w_giou=2.0, w_ce_b=1.0, pos_weight=10.0, eos_coef=0.1,
                   focal_alpha=0.25, focal_gamma=2.0, label_smoothing=0.1)


def cxcywh_to_xyxy(b):
    cx, cy, w, h = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)


def giou_matrix(a, t):
    """generalized_box_iou (utils/box_ops.py:83-108) of xyxy boxes a [n,4], t [m,4] -> [n,m]."""
    area1 = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area2 = (t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1])
    lt = np.maximum(a[:, None, :2], t[None, :, :2]); rb = np.minimum(a[:, None, 2:], t[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    union = area1[:, None] + area2[None] - inter
    iou = inter / union
    lt = np.minimum(a[:, None, :2], t[None, :, :2]); rb = np.maximum(a[:, None, 2:], t[None, :, 2:])
    wh = np.clip(rb - lt, 0, None)
    area = wh[..., 0] * wh[..., 1]
    return iou - (area - union) / area


def softmax(z):
    e = np.exp(z - z.max(-1, keepdims=True))
    return e / e.sum(-1, keepdims=True)


def cost_matrix(pred_boxes, pred_logits_b, tgt_boxes, cfg):
    """matcher.py:58-71 for ONE clip: pred_boxes [nq,4], pred_logits_b [nq,3], tgt_boxes [n,4] -> [nq,n]."""
    l1 = np.abs(pred_boxes[:, None, :] - tgt_boxes[None]).sum(-1)
    g = -giou_matrix(cxcywh_to_xyxy(pred_boxes), cxcywh_to_xyxy(tgt_boxes))
    pc = -softmax(pred_logits_b)[:, 1:2]
    return cfg["cost_bbox"] * l1 + cfg["cost_giou"] * g + cfg["cost_class"] * pc


def match_ava(pred_boxes, pred_logits_b, tgt_boxes, n_tgt, cfg, dtype=np.float32):
    """-> match [B,nq] int32: target index or -1.  The cost is formed in `dtype` (the reference forms it in fp32 on the device)
    and solved in fp64 by scipy, as in matcher.py:72-77."""
    B, nq = pred_boxes.shape[:2]
    match = -np.ones((B, nq), dtype=np.int32)
    for b in range(B):
        n = int(n_tgt[b])
        if n == 0:
            continue
        C = cost_matrix(pred_boxes[b].astype(dtype), pred_logits_b[b].astype(dtype), tgt_boxes[b, :n].astype(dtype), cfg)
        qi, tj = linear_sum_assignment(np.asarray(C, dtype=np.float64))
        match[b, qi] = tj
    return match


def focal(x, t, w, alpha, gamma):
    p = 1.0 / (1.0 + np.exp(-x))
    ce = np.maximum(x, 0) - x * t + np.log1p(np.exp(-np.abs(x)))
    ce = ce * w
    pt = p * t + (1 - p) * (1 - t)
    loss = ce * (1 - pt) ** gamma
    if alpha >= 0:
        loss = (alpha * t + (1 - alpha) * (1 - t)) * loss
    return loss


def criterion_ava(pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, n_tgt, cfg=None, match=None):
    """-> dict(loss_ce, loss_bbox, loss_giou, loss_ce_b, total, class_error, match).  Arrays as in include/cqvad.h
    (cqvad_criterion_ava): padded targets [B,maxT,4] / [B,maxT,K] with n_tgt valid rows."""
    cfg = dict(DEFAULT_CFG, **(cfg or {}))
    pl, pb, plb = (np.asarray(a, dtype=np.float64) for a in (pred_logits, pred_boxes, pred_logits_b))
    B, nq, K = pl.shape
    if match is None:
        match = match_ava(np.asarray(pred_boxes), np.asarray(pred_logits_b), np.asarray(tgt_boxes), n_tgt, cfg)
    sm = cfg["label_smoothing"]
    tl, fl = (1 - sm) * 1 + 0.5 * sm, 0.5 * sm
    T = np.full((B, nq, K), fl); Wt = np.ones((B, nq, 1))
    n_p, l1, gi, pairs, hits = 0.0, 0.0, 0.0, 0, 0
    y = np.full((B, nq), 2, dtype=np.int64)
    for b in range(B):
        for q in range(nq):
            j = int(match[b, q])
            if j < 0:
                continue
            lab = np.asarray(tgt_labels[b, j], dtype=np.float64)
            n_p += lab.sum()
            T[b, q] = np.where(lab == 0, fl, np.where(lab == 1, tl, lab)) if sm else lab
            Wt[b, q] = cfg["pos_weight"]; y[b, q] = 1
            tb = np.asarray(tgt_boxes[b, j], dtype=np.float64)
            l1 += np.abs(pb[b, q] - tb).sum()
            gi += 1.0 - giou_matrix(cxcywh_to_xyxy(pb[b, q][None]), cxcywh_to_xyxy(tb[None]))[0, 0]
            pairs += 1
            if sm:
                hits += 1                        # accuracy_sigmoid sees the smoothed (all non-zero) targets: criterion.py:68,76-77,103
            else:
                nl = int((lab != 0).sum())
                top = np.argsort(-pl[b, q], kind="stable")[:nl]
                hits += int(set(top.tolist()) == set(np.nonzero(lab)[0].tolist()))
    n_p = max(n_p, 1.0)
    num_boxes = float(np.sum(n_tgt))
    loss_ce = focal(pl, T, Wt, cfg["focal_alpha"], cfg["focal_gamma"]).mean(-1).sum() / n_p
    lsm = plb - plb.max(-1, keepdims=True); lsm = lsm - np.log(np.exp(lsm).sum(-1, keepdims=True))
    wy = np.where(y == 2, cfg["eos_coef"], 1.0)
    nll = -np.take_along_axis(lsm, y[..., None], -1)[..., 0]
    loss_ce_b = (wy * nll).sum() / wy.sum()
    out = dict(loss_ce=loss_ce, loss_bbox=l1 / num_boxes, loss_giou=gi / num_boxes, loss_ce_b=loss_ce_b,
               class_error=100.0 - hits * (100.0 / pairs) if pairs else 100.0, match=match)
    out["total"] = (cfg["w_ce"] * out["loss_ce"] + cfg["w_bbox"] * out["loss_bbox"] + cfg["w_giou"] * out["loss_giou"]
                    + cfg["w_ce_b"] * out["loss_ce_b"])
    return out


def postprocess_ava(pred_logits, pred_boxes, pred_logits_b, target_sizes):
    """criterion.py:740-773 -> detections [B,nq,K+5] = [scores | boxes xyxy in pixels | person probability]."""
    pl, pb, plb = (np.asarray(a, dtype=np.float64) for a in (pred_logits, pred_boxes, pred_logits_b))
    ts = np.asarray(target_sizes, dtype=np.float64)
    scale = np.stack([ts[:, 1], ts[:, 0], ts[:, 1], ts[:, 0]], 1)[:, None, :]
    return np.concatenate([1.0 / (1.0 + np.exp(-pl)), cxcywh_to_xyxy(pb) * scale, softmax(plb)[..., 1:2]], -1)


def postprocess_ucf(pred_logits, pred_boxes, pred_logits_b, target_sizes):
    """PostProcessUCF / PostProcessJHMDB (criterion.py:775-846): scores gated by the person probability through inverse_sigmoid
    (utils/misc.py:530-534, eps 1e-5) and sigmoid."""
    det = postprocess_ava(pred_logits, pred_boxes, pred_logits_b, target_sizes)
    K = np.asarray(pred_logits).shape[-1]
    x = np.clip(det[..., :K] * det[..., K + 4:K + 5], 0.0, 1.0)
    y = np.log(np.maximum(x, 1e-5) / np.maximum(1 - x, 1e-5))
    det[..., :K] = 1.0 / (1.0 + np.exp(-y))
    return det
