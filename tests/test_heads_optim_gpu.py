"""Training-step pieces outside the decoder: DETR heads (models/model.py:191-236) fwd/bwd against a fixture made with the reference's
MLP / inverse_sigmoid + autograd, the Philox dropout of the class tokens (statistics + forward/backward mask agreement), and the
fused clip_grad_norm_ + AdamW step (train.py:83,158-167) against torch's own implementation of the two on the same device."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err

pytestmark = pytest.mark.gpu


def _heads(dev, W):
    from class_query_vad_b200 import DETRHeads
    h = DETRHeads()
    h.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    return h.to(dev)


def test_heads_forward_backward_match_reference():
    from oracle.make_golden_heads import make_case
    g = load_golden("heads")
    hs, cls_hs, refs, W, lw = make_case()
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    heads = _heads(dev, W).eval()          # eval: Dropout(0.5) is the identity, as in the fixture
    x = {k: t(v).requires_grad_(True) for k, v in (("hs", hs), ("cls_hs", cls_hs), ("refs", refs))}
    out = heads(x["hs"], x["cls_hs"], x["refs"])
    pl = torch.stack([a["pred_logits"] for a in out["aux_outputs"]] + [out["pred_logits"]])
    pb = torch.stack([a["pred_boxes"] for a in out["aux_outputs"]] + [out["pred_boxes"]])
    plb = torch.stack([a["pred_logits_b"] for a in out["aux_outputs"]] + [out["pred_logits_b"]])
    for got, key in ((pl, "pred_logits"), (pb, "pred_boxes"), (plb, "pred_logits_b")):
        assert rel_err(got.detach().cpu().numpy(), g[key]) < 1e-5, key
    loss = (pl * t(lw["w_logits"])).sum() + (pb * t(lw["w_boxes"])).sum() + (plb * t(lw["w_logits_b"])).sum()
    loss.backward()
    for k, key in (("hs", "g_hs"), ("cls_hs", "g_cls_hs"), ("refs", "g_refs")):
        assert rel_err(x[k].grad.cpu().numpy(), g[key]) < 1e-4, key
    for n, p in heads.named_parameters():
        assert rel_err(p.grad.cpu().numpy(), g["g." + n]) < 1e-4, n


def test_heads_dropout_mask_statistics_and_forward_backward_agreement():
    from oracle.make_golden_heads import make_case
    _, _, _, W, _ = make_case()
    dev = torch.device("cuda:0")
    heads = _heads(dev, W).train()
    Lr, BT, nq, K = 1, 8, 15, 80
    gen = torch.Generator(device="cpu").manual_seed(0)
    hs = torch.randn((Lr, BT, nq, 256), generator=gen).to(dev)
    cls_hs = torch.randn((Lr, BT, nq, K, 256), generator=gen).to(dev).requires_grad_(True)
    refs = torch.rand((Lr, BT, nq, 4), generator=gen).to(dev)
    out = heads(hs, cls_hs, refs)
    out["pred_logits"].sum().backward()
    gmask = cls_hs.grad * 256.0 * 0.5                     # = keep mask (1 / ((1 - p) * 256) on kept elements)
    keep = gmask.round()
    assert float((gmask - keep).abs().max()) < 1e-6 and set(keep.unique().tolist()) <= {0.0, 1.0}
    n = keep.numel()
    rate = float(keep.mean())
    assert abs(rate - 0.5) < 5 * 0.5 / np.sqrt(n), rate                       # keep rate 1 - p within 5 sigma (n = 2.5 M)
    # no structure along the channel or the row axis
    assert float((keep.mean(dim=-1) - 0.5).abs().max()) < 6 * 0.5 / np.sqrt(256)
    assert float((keep.flatten(0, -2).mean(dim=0) - 0.5).abs().max()) < 6 * 0.5 / np.sqrt(n / 256)
    # the forward used the same mask: mean(mask * x / (1 - p))
    ref = (keep * cls_hs.detach() * 2.0).mean(-1)[-1]
    assert rel_err(out["pred_logits"].detach().cpu().numpy(), ref.cpu().numpy()) < 1e-5
    # a second call draws a different mask; eval mode is exact
    out2 = heads(hs, cls_hs.detach(), refs)
    assert float((out2["pred_logits"] - out["pred_logits"]).detach().abs().max()) > 1e-3
    heads.eval()
    assert rel_err(heads(hs, cls_hs.detach(), refs)["pred_logits"].detach().cpu().numpy(), cls_hs.detach().mean(-1)[-1].cpu().numpy()) < 1e-5


@pytest.mark.parametrize("max_norm", [1.0, 1e9])
def test_flat_adamw_matches_torch_adamw_with_clipping(max_norm):
    from class_query_vad_b200 import FlatAdamW
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(1)
    shapes = [(256, 256), (256,), (3, 256), (7,), (1,), (1024, 256), (5, 3, 3)]
    mine = [torch.nn.Parameter(torch.randn(s, generator=gen).to(dev)) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in mine]
    opt = FlatAdamW([(f"p{i}", p) for i, p in enumerate(mine)], lr=1e-2, max_norm=max_norm)
    topt = torch.optim.AdamW(ref, lr=1e-2)
    for step in range(4):
        grads = [torch.randn(s, generator=gen).to(dev) * (10.0 if step % 2 else 0.01) for s in shapes]
        for p, r, g in zip(mine, ref, grads):
            p.grad.copy_(g)                    # persistent views of the flat gradient buffer
            r.grad = g.clone()
        tn = torch.nn.utils.clip_grad_norm_(ref, max_norm=max_norm)
        topt.step()
        norm = opt.step()
        assert abs(float(norm) - float(tn)) < 1e-5 * float(tn)
        for i, (p, r) in enumerate(zip(mine, ref)):
            assert float((p.detach() - r.detach()).abs().max()) < 2e-6 * max(1.0, float(r.detach().abs().max())), (step, i)
            assert float(p.grad.abs().max()) == 0.0          # zero_grad folded into the step


def test_flat_adamw_grad_scale_is_the_all_reduce_average():
    from class_query_vad_b200 import FlatAdamW
    dev = torch.device("cuda:0")
    a = torch.nn.Parameter(torch.ones(1000, device=dev)); b = torch.nn.Parameter(torch.ones(1000, device=dev))
    oa, ob = FlatAdamW([("a", a)], lr=1e-2, max_norm=1.0), FlatAdamW([("b", b)], lr=1e-2, max_norm=1.0)
    g = torch.linspace(-1, 1, 1000, device=dev)
    a.grad.copy_(g * 8.0); b.grad.copy_(g)
    oa.step(grad_scale=1.0 / 8.0); ob.step()
    assert float((a.detach() - b.detach()).abs().max()) < 1e-7


def test_flat_adamw_zero_gradient_step_only_decays():
    """Property: with zero gradients AdamW moves a parameter by the decoupled weight decay alone, p <- p (1 - lr wd) (moments stay 0),
    and by nothing at all with weight_decay = 0."""
    from class_query_vad_b200 import FlatAdamW
    dev = torch.device("cuda:0")
    a = torch.nn.Parameter(torch.linspace(-2, 2, 1003, device=dev)); b = torch.nn.Parameter(a.detach().clone())
    ref = a.detach().clone()
    FlatAdamW([("a", a)], lr=1e-2, weight_decay=0.1).step()
    FlatAdamW([("b", b)], lr=1e-2, weight_decay=0.0).step()
    assert float((a.detach() - ref * (1 - 1e-2 * 0.1)).abs().max()) < 1e-6
    assert torch.equal(b.detach(), ref)
