"""Loss / matcher / post-process: oracle and CUDA path against fixtures generated from the reference's
own HungarianMatcherAVA / SetCriterionAVA / PostProcessAVA (oracle/make_golden_criterion.py)."""
import numpy as np
import pytest

from oracle import criterion_np
from helpers import load_golden, rel_err, TOL_FP32

CASES = ["criterion_ava", "criterion_nosmooth", "criterion_more_targets", "criterion_small"]
LOSS_KEYS = ("loss_ce", "loss_bbox", "loss_giou", "loss_ce_b", "total", "class_error")


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_criterion(name):
    g = load_golden(name)
    cfg = {"label_smoothing": float(g["smooth"])}
    out = criterion_np.criterion_ava(g["pred_logits"], g["pred_boxes"], g["pred_logits_b"], g["tgt_boxes"], g["tgt_labels"], g["n_tgt"], cfg)
    assert np.array_equal(out["match"], g["match"])                    # assignment: exact
    for k, ref in zip(LOSS_KEYS, g["losses"]):
        assert abs(out[k] - ref) <= 2e-6 * max(1.0, abs(ref)), (k, out[k], ref)
    det = criterion_np.postprocess_ava(g["pred_logits"], g["pred_boxes"], g["pred_logits_b"], g["sizes"])
    assert rel_err(det, g["det"]) < 1e-6
    det_u = criterion_np.postprocess_ucf(g["pred_logits"], g["pred_boxes"], g["pred_logits_b"], g["sizes"])
    assert rel_err(det_u, g["det_ucf"]) < 1e-6


def _gpu_inputs(g):
    import torch
    dev = torch.device("cuda:0")
    return {k: torch.from_numpy(np.ascontiguousarray(g[k])).to(dev) for k in
            ("pred_logits", "pred_boxes", "pred_logits_b", "tgt_boxes", "tgt_labels", "n_tgt", "sizes")}


def _modules(g):
    from class_query_vad_b200.modules.criterion import HungarianMatcherAVA, SetCriterionAVA, PostProcessAVA
    c = criterion_np.DEFAULT_CFG
    matcher = HungarianMatcherAVA(cost_class=c["cost_class"], cost_bbox=c["cost_bbox"], cost_giou=c["cost_giou"])
    wd = {"loss_ce": c["w_ce"], "loss_bbox": c["w_bbox"], "loss_giou": c["w_giou"], "loss_ce_b": c["w_ce_b"]}
    K = g["pred_logits"].shape[-1]
    crit = SetCriterionAVA(c["pos_weight"], K, num_queries=g["pred_logits"].shape[1], matcher=matcher, weight_dict=wd,
                           eos_coef=c["eos_coef"], losses=["labels", "boxes"], data_file="ava")
    crit.label_smoothing_alpha = float(g["smooth"])
    return matcher, crit, PostProcessAVA()


def _targets(g, t):
    import torch
    out = []
    for b, n in enumerate(g["n_tgt"]):
        n = int(n)
        out.append({"boxes": torch.cat([torch.zeros(n, 1, device=t["tgt_boxes"].device), t["tgt_boxes"][b, :n]], 1),
                    "labels": t["tgt_labels"][b, :n]})
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_device_criterion_matches_reference(name):
    """Assignment bit-exact; losses, weighted total and its gradient at the fp32 tolerance (1e-3; measured ~1e-6)."""
    import torch
    g = load_golden(name)
    t = _gpu_inputs(g)
    matcher, crit, post = _modules(g)
    outputs = {k: t[k] for k in ("pred_logits", "pred_boxes", "pred_logits_b")}
    losses, match, grads = crit.total_and_grads(outputs, t["tgt_boxes"], t["tgt_labels"], t["n_tgt"])
    assert np.array_equal(match.cpu().numpy(), g["match"])
    L = losses.cpu().numpy()
    for i, (k, ref) in enumerate(zip(LOSS_KEYS, g["losses"])):
        assert abs(L[i] - ref) <= 1e-5 * max(1.0, abs(ref)), (k, L[i], ref)
    assert int(L[6]) == int((g["match"] >= 0).sum()) and L[7] == float(g["n_tgt"].sum())
    for got, key in zip(grads, ("g_logits", "g_boxes", "g_logits_b")):
        assert rel_err(got.cpu().numpy(), g[key]) < 1e-4, key
    # reference-format entry points: matcher indices and the autograd route of the loss dict
    idx = matcher(outputs, _targets(g, t))
    for b, (qi, tj) in enumerate(idx):
        ref_q = np.nonzero(g["match"][b] >= 0)[0]
        assert np.array_equal(qi.numpy(), ref_q) and np.array_equal(tj.numpy(), g["match"][b][ref_q])
    req = {k: v.clone().requires_grad_(True) for k, v in outputs.items()}
    ld = crit(req, _targets(g, t))
    total = sum(ld[k] * w for k, w in crit.weight_dict.items())          # train.py:148
    total.backward()
    assert abs(float(total) - g["losses"][4]) <= 1e-5 * abs(g["losses"][4])
    for k, key in (("pred_logits", "g_logits"), ("pred_boxes", "g_boxes"), ("pred_logits_b", "g_logits_b")):
        assert rel_err(req[k].grad.cpu().numpy(), g[key]) < 1e-4, key
    scores, boxes, person = post(outputs, t["sizes"])
    det = np.concatenate([scores, boxes, person], -1)
    assert rel_err(det, g["det"]) < 1e-5
    from class_query_vad_b200 import PostProcessUCF, PostProcessJHMDB
    for cls in (PostProcessUCF, PostProcessJHMDB):
        s_u, b_u, p_u = cls()(outputs, t["sizes"])
        K = g["pred_logits"].shape[-1]
        assert np.abs(s_u - g["det_ucf"][..., :K]).max() < 1e-6            # probabilities: absolute
        assert rel_err(np.concatenate([b_u, p_u], -1), g["det_ucf"][..., K:]) < 1e-5


@pytest.mark.gpu
def test_device_matcher_is_exact_on_random_costs():
    """Size-independent property at the ABI limits (64 queries x 64 targets): the device assignment has the same total cost as
    scipy's on the same fp32 cost matrix, and is a valid one-to-one assignment."""
    import torch
    from scipy.optimize import linear_sum_assignment
    from class_query_vad_b200.modules.criterion import HungarianMatcherAVA
    rs = np.random.RandomState(0)
    B, nq, K = 8, 64, 4
    n_tgt = np.array([64, 1, 0, 33, 64, 17, 5, 48], dtype=np.int32)
    pb = np.concatenate([rs.uniform(0.2, 0.8, (B, nq, 2)), rs.uniform(0.05, 0.5, (B, nq, 2))], -1).astype(np.float32)
    tb = np.concatenate([rs.uniform(0.2, 0.8, (B, 64, 2)), rs.uniform(0.05, 0.5, (B, 64, 2))], -1).astype(np.float32)
    plb = rs.standard_normal((B, nq, 3)).astype(np.float32)
    dev = torch.device("cuda:0")
    m = HungarianMatcherAVA(cost_class=12, cost_bbox=5, cost_giou=2)
    out = {"pred_logits": torch.zeros(B, nq, K, device=dev), "pred_boxes": torch.from_numpy(pb).to(dev),
           "pred_logits_b": torch.from_numpy(plb).to(dev)}
    match = m.match_dense(out, torch.from_numpy(tb).to(dev), torch.zeros(B, 64, K, device=dev), torch.from_numpy(n_tgt).to(dev)).cpu().numpy()
    cfg = criterion_np.DEFAULT_CFG
    for b in range(B):
        n = int(n_tgt[b])
        q = np.nonzero(match[b] >= 0)[0]
        assert len(q) == min(n, nq) and len(set(match[b, q].tolist())) == len(q)
        if n == 0:
            continue
        C = criterion_np.cost_matrix(pb[b], plb[b], tb[b, :n], cfg).astype(np.float32).astype(np.float64)
        ri, ci = linear_sum_assignment(C)
        assert abs(C[q, match[b, q]].sum() - C[ri, ci].sum()) < 1e-4
