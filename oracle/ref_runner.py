"""BASELINE INFRASTRUCTURE ONLY -- runs the UNMODIFIED reference decoder (models/detr/dab_transformer.py TransformerDecoder +
the 30 lines of DETR head logic of models/model.py:191-221 restated, because `models.model` itself cannot be imported:
yacs / timm / VideoMamba) from the install under baseline/_ref (oracle/install_ref.py, oracle/ref_import.py).

Used by bench.py only:  `--impl reference` / `cpu_baseline` (the reference on the box's host cores) and the
`reference_gpu_eager` extra key (the reference's own eager PyTorch on the same B200: fp32 with TF32 off, and bf16 autocast --
BASELINE.md section 5.1).  Never imported by the product package.
"""
import os

import numpy as np
import torch

from . import synth
from .ref_import import import_reference, available  # noqa: F401


def build(cfg_name_or_dict, seed=0, device="cpu", train=False):
    """-> (decoder module, cfg dict, heads dict of tensors).  eval(): dropout = identity (parity semantics); train=True keeps the
    reference's dropout (p = 0.1) active, as in its own training loop (train.py:126-182)."""
    from .make_golden import build_reference_decoder
    ref = import_reference()
    cfg = dict(synth.CONFIGS[cfg_name_or_dict]) if isinstance(cfg_name_or_dict, str) else dict(cfg_name_or_dict)
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=seed)
    dec = build_reference_decoder(ref, cfg, W).to(device)
    if train:
        dec.train()
    heads = {k: torch.from_numpy(W[k]).to(device) for k in ("heads.class_embed_b.weight", "heads.class_embed_b.bias")}
    return dec, cfg, heads


def _inverse_sigmoid(x, eps=1e-5):           # utils/misc.py:530-534
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def heads_forward(dec, heads, hs, cls_hs, refs):
    """models/model.py:191-221 (eval): class_embed_b, bbox_embed + inverse_sigmoid(reference) -> sigmoid, cls_hs.mean(-1)."""
    logits_b = torch.nn.functional.linear(hs, heads["heads.class_embed_b.weight"], heads["heads.class_embed_b.bias"])
    tmp = dec.bbox_embed(hs)
    tmp = tmp[..., :4] + _inverse_sigmoid(refs)
    return cls_hs.mean(-1), tmp.sigmoid(), logits_b


def make_inputs(cfg_name, B, seed, device):
    inp = synth.make_decoder_inputs(cfg_name, B, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
    return {k: t(inp[k]) for k in ("tgt", "memory", "mask", "pos", "refpoints_unsigmoid")}, inp["orig_res"]


def step_fn(cfg_name, mode, B=1, device="cpu", autocast_bf16=False, seed=0):
    """One step of the benchmark workload on the unmodified reference: mode 'infer' = decoder forward + heads under no_grad;
    mode 'train' = forward + loss = sum(w*out) + autograd backward (gradients of every parameter, memory, tgt, refpoints)."""
    dec, cfg, heads = build(cfg_name, seed=seed, device=device)
    x, orig_res = make_inputs(cfg_name, B, seed, device)
    lw = synth.make_loss_weights(cfg, B, seed=1)
    w_hs, w_cls, w_refs = (torch.from_numpy(lw[k]).to(device) for k in ("w_hs", "w_cls", "w_refs"))
    dev_type = "cuda" if str(device).startswith("cuda") else "cpu"

    def fwd(tgt, memory, refu):
        with torch.autocast(dev_type, dtype=torch.bfloat16, enabled=autocast_bf16):
            return dec(tgt, memory, memory_key_padding_mask=x["mask"], pos=x["pos"], refpoints_unsigmoid=refu, orig_res=orig_res)

    if mode == "infer":
        def step():
            with torch.no_grad():
                hs, cls_hs, refs = fwd(x["tgt"], x["memory"], x["refpoints_unsigmoid"])
                return heads_forward(dec, heads, hs.float(), cls_hs.float(), refs.float())
        return step

    def step():
        for p in dec.parameters():
            p.grad = None
        tgt = x["tgt"].clone().requires_grad_(True)
        memory = x["memory"].clone().requires_grad_(True)
        refu = x["refpoints_unsigmoid"].clone().requires_grad_(True)
        hs, cls_hs, refs = fwd(tgt, memory, refu)
        loss = (w_hs * hs.float()).sum() + (w_cls * cls_hs.float()).sum() + (w_refs * refs.float()).sum()
        loss.backward()
        return loss
    return step


def cpu_threads():
    n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return n
