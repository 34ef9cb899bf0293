"""TEST INFRASTRUCTURE ONLY.  tests/golden/heads.npz: the head lines of the reference DETR.forward (models/model.py:192-199,
219-221, restated verbatim below because they live inside forward() of the full model) evaluated with the reference's own MLP
(models/detr/dab_transformer.py:36-48), inverse_sigmoid (utils/misc.py:530-534) and nn.Linear, plus torch autograd gradients of
a fixed linear loss.  Dropout is the identity (eval): the p = 0.5 mask of training has no reference RNG stream to match and is
tested statistically.  Run in the build container only:   python -m oracle.make_golden_heads"""
import importlib
import os
import numpy as np
import torch

from .ref_import import import_reference
from .make_golden import GOLD


def make_case(seed=0, Lr=2, BT=2, nq=3, K=5):
    rs = np.random.RandomState(8000 + seed)
    hs = rs.standard_normal((Lr, BT, nq, 256)).astype(np.float32)
    cls_hs = rs.standard_normal((Lr, BT, nq, K, 256)).astype(np.float32)
    refs = rs.uniform(0.02, 0.98, (Lr, BT, nq, 4)).astype(np.float32)
    refs[0, 0, 0] = [0.0, 1.0, 5e-6, 0.5]                   # clamp branches of inverse_sigmoid
    W = {"bbox_embed.layers.0.weight": rs.standard_normal((256, 256)) / 16, "bbox_embed.layers.0.bias": 0.1 * rs.standard_normal(256),
         "bbox_embed.layers.1.weight": rs.standard_normal((256, 256)) / 16, "bbox_embed.layers.1.bias": 0.1 * rs.standard_normal(256),
         "bbox_embed.layers.2.weight": rs.standard_normal((4, 256)) / 16, "bbox_embed.layers.2.bias": 0.1 * rs.standard_normal(4),
         "class_embed_b.weight": rs.standard_normal((3, 256)) / 16, "class_embed_b.bias": 0.1 * rs.standard_normal(3)}
    W = {k: v.astype(np.float32) for k, v in W.items()}
    lw = {"w_logits": rs.standard_normal((Lr, BT, nq, K)).astype(np.float32), "w_boxes": rs.standard_normal((Lr, BT, nq, 4)).astype(np.float32),
          "w_logits_b": rs.standard_normal((Lr, BT, nq, 3)).astype(np.float32)}
    return hs, cls_hs, refs, W, lw


def main():
    ref = import_reference()
    misc = importlib.import_module("utils.misc")
    hs, cls_hs, refs, W, lw = make_case()
    bbox_embed = ref.MLP(256, 256, 4, 3)
    class_embed_b = torch.nn.Linear(256, 3)
    bbox_embed.load_state_dict({k[len("bbox_embed."):]: torch.from_numpy(v) for k, v in W.items() if k.startswith("bbox_embed.")})
    class_embed_b.load_state_dict({k[len("class_embed_b."):]: torch.from_numpy(v) for k, v in W.items() if k.startswith("class_embed_b.")})
    t = {k: torch.from_numpy(v).requires_grad_(True) for k, v in (("hs", hs), ("cls_hs", cls_hs), ("refs", refs))}
    outputs_class_b = class_embed_b(t["hs"])                                   # model.py:192
    reference_before_sigmoid = misc.inverse_sigmoid(t["refs"])                 # :196
    tmp = bbox_embed(t["hs"])                                                  # :197
    tmp[..., :4] += reference_before_sigmoid                                   # :198
    outputs_coord = tmp.sigmoid()                                              # :199
    outputs_class = t["cls_hs"].mean(dim=-1)                                   # :219-221 with Dropout in eval mode
    loss = (outputs_class * torch.from_numpy(lw["w_logits"])).sum() + (outputs_coord * torch.from_numpy(lw["w_boxes"])).sum() + \
        (outputs_class_b * torch.from_numpy(lw["w_logits_b"])).sum()
    loss.backward()
    out = dict(pred_logits=outputs_class.detach().numpy(), pred_boxes=outputs_coord.detach().numpy(),
               pred_logits_b=outputs_class_b.detach().numpy(), g_hs=t["hs"].grad.numpy(), g_cls_hs=t["cls_hs"].grad.numpy(),
               g_refs=t["refs"].grad.numpy())
    for n, p in list(bbox_embed.named_parameters()):
        out["g.bbox_embed." + n] = p.grad.numpy()
    for n, p in class_embed_b.named_parameters():
        out["g.class_embed_b." + n] = p.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, "heads.npz"), **out)
    print("heads", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
