"""TEST INFRASTRUCTURE ONLY.  tests/golden/criterion_*.npz: the reference's own HungarianMatcherAVA (models/detr/matcher.py:13-78),
SetCriterionAVA (models/detr/criterion.py:17-224), the weighted total of train.py:148 with its autograd gradient with respect
to the three prediction tensors, and PostProcessAVA (criterion.py:740-773), run unmodified on synthetic predictions / targets.
Run in the build container only:   python -m oracle.make_golden_criterion"""
import importlib
import os
import numpy as np
import torch

from .ref_import import import_reference
from .make_golden import GOLD
from .criterion_np import DEFAULT_CFG, match_ava

# (fixture, nq, K, targets per clip, label smoothing, seed)
CASES = [
    ("criterion_ava", 15, 80, [3, 0, 1, 7, 2, 15], 0.1, 0),         # AVA22_ViT-B: 15 queries, 80 classes; one clip saturates the queries
    ("criterion_nosmooth", 15, 80, [2, 4, 1], 0.0, 1),               # class_error is only meaningful without smoothing
    ("criterion_more_targets", 5, 21, [8, 2, 6], 0.1, 2),            # more targets than queries (rectangular assignment, transposed)
    ("criterion_small", 3, 4, [1, 2], 0.1, 3),
]


def make_case(nq, K, n_tgt, seed):
    rs = np.random.RandomState(7000 + seed)
    B, maxT = len(n_tgt), max(max(n_tgt), 1)
    pred_logits = (1.5 * rs.standard_normal((B, nq, K))).astype(np.float32)
    pred_boxes = np.concatenate([rs.uniform(0.2, 0.8, (B, nq, 2)), rs.uniform(0.05, 0.5, (B, nq, 2))], -1).astype(np.float32)
    pred_logits_b = (1.5 * rs.standard_normal((B, nq, 3))).astype(np.float32)
    tgt_boxes = np.concatenate([rs.uniform(0.2, 0.8, (B, maxT, 2)), rs.uniform(0.05, 0.5, (B, maxT, 2))], -1).astype(np.float32)
    tgt_labels = np.zeros((B, maxT, K), dtype=np.float32)
    for b in range(B):
        for j in range(maxT):
            tgt_labels[b, j, rs.choice(K, size=rs.randint(1, min(4, K) + 1), replace=False)] = 1.0
    # a prediction placed exactly on a target (max/min ties of the GIoU gradient) and a disjoint pair (zero intersection)
    pred_boxes[0, 0] = tgt_boxes[0, 0]
    pred_boxes[0, 1] = np.array([0.1, 0.1, 0.05, 0.05], dtype=np.float32)
    sizes = np.stack([rs.randint(200, 400, B), rs.randint(300, 500, B)], 1).astype(np.float32)     # (h, w)
    # every other matched row predicts its target's label set (so that class_error is neither 0 nor 100 without smoothing); the
    # assignment does not depend on pred_logits (matcher.py:68-71)
    m = match_ava(pred_boxes, pred_logits_b, tgt_boxes, n_tgt, DEFAULT_CFG)
    for b, q in zip(*np.nonzero(m >= 0)):
        if q % 2 == 0:
            pred_logits[b, q] += 8.0 * tgt_labels[b, m[b, q]]
    return pred_logits, pred_boxes, pred_logits_b, tgt_boxes, tgt_labels, np.asarray(n_tgt, dtype=np.int32), sizes


def main():
    import_reference()
    crit_mod = importlib.import_module("models.detr.criterion")
    match_mod = importlib.import_module("models.detr.matcher")
    c = DEFAULT_CFG
    for name, nq, K, n_tgt, smooth, seed in CASES:
        pl, pb, plb, tb, tlab, nt, sizes = make_case(nq, K, n_tgt, seed)
        matcher = match_mod.HungarianMatcherAVA(cost_class=c["cost_class"], cost_bbox=c["cost_bbox"], cost_giou=c["cost_giou"])
        weight_dict = {"loss_ce": c["w_ce"], "loss_bbox": c["w_bbox"], "loss_giou": c["w_giou"], "loss_ce_b": c["w_ce_b"]}
        crit = crit_mod.SetCriterionAVA(c["pos_weight"], K, num_queries=nq, matcher=matcher, weight_dict=weight_dict,
                                        eos_coef=c["eos_coef"], losses=["labels", "boxes"], data_file="ava")
        crit.label_smoothing_alpha = smooth          # the constructor hard-codes 0.1 (criterion.py:48)
        t = [torch.from_numpy(a).clone().requires_grad_(True) for a in (pl, pb, plb)]
        outputs = {"pred_logits": t[0], "pred_boxes": t[1], "pred_logits_b": t[2]}
        # reference target format: boxes [n,5] (column 0 = key-frame id, dropped by matcher.py:60 / criterion.py:128), labels [n,K]
        targets = [{"boxes": torch.from_numpy(np.concatenate([np.zeros((int(n), 1), np.float32), tb[b, :n]], 1)),
                    "labels": torch.from_numpy(tlab[b, :n].copy())} for b, n in enumerate(nt)]
        indices = matcher({k: v.detach() for k, v in outputs.items()}, targets)
        match = -np.ones((len(nt), nq), dtype=np.int32)
        for b, (qi, tj) in enumerate(indices):
            match[b, qi.numpy()] = tj.numpy()
        loss_dict = crit(outputs, targets)
        total = sum(loss_dict[k] * w for k, w in crit.weight_dict.items())           # train.py:148
        total.backward()
        scores, boxes, person = crit_mod.PostProcessAVA()({k: v.detach() for k, v in outputs.items()}, torch.from_numpy(sizes))
        s_u, b_u, p_u = crit_mod.PostProcessUCF()({k: v.detach() for k, v in outputs.items()}, torch.from_numpy(sizes))
        s_j, b_j, p_j = crit_mod.PostProcessJHMDB()({k: v.detach() for k, v in outputs.items()}, torch.from_numpy(sizes))
        assert np.array_equal(s_u, s_j) and np.array_equal(b_u, b_j)
        ce = loss_dict["class_error"]
        np.savez_compressed(
            os.path.join(GOLD, name + ".npz"), pred_logits=pl, pred_boxes=pb, pred_logits_b=plb, tgt_boxes=tb, tgt_labels=tlab,
            n_tgt=nt, sizes=sizes, smooth=np.float32(smooth), match=match,
            losses=np.array([float(loss_dict[k]) for k in ("loss_ce", "loss_bbox", "loss_giou", "loss_ce_b")] + [float(total), float(ce)]),
            g_logits=t[0].grad.numpy(), g_boxes=t[1].grad.numpy(), g_logits_b=t[2].grad.numpy(),
            det=np.concatenate([scores, boxes, person], -1).astype(np.float32),
            det_ucf=np.concatenate([s_u, b_u, p_u], -1).astype(np.float32))
        print(name, {k: float(v) for k, v in loss_dict.items()}, "total", float(total), "pairs", int((match >= 0).sum()))


if __name__ == "__main__":
    main()
