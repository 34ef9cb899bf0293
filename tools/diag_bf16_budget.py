"""bf16 error budget of the decoder training step (run on the GPU box): per-tensor relative-L2 error of outputs and gradients
for   (A) the bf16 path vs the reference-autograd fixture,
      (B) the fp32 path with the MATRICES rounded to bf16 vs the fixture        = weight-quantisation error alone,
      (C) the bf16 path vs (B)                                                  = arithmetic / storage error alone.
usage: python tools/diag_bf16_budget.py grad_ava_vitb_b1_l2 [grad_ava_vitb_b2_l6 ...]   (TOPN=15 rows per table)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from helpers import load_golden, case_from_golden
from test_train_gpu import run_train, grad_errors, ZERO_BIAS, ZERO_TGT0
from class_query_vad_b200 import _lib

TOPN = int(os.environ.get("TOPN", "15"))


def round_matrices(W, layers):
    lib = _lib.lib()
    kinds = {lib.cqvad_decoder_weight_name(i, layers).decode(): lib.cqvad_decoder_weight_kind(i, layers)
             for i in range(lib.cqvad_decoder_num_weights(layers))}
    out = {}
    for k, v in W.items():
        if kinds.get(k, 1) == 0:
            out[k] = torch.from_numpy(v).bfloat16().float().numpy()
        else:
            out[k] = v
    return out


def summary(tag, l2):
    v = np.array(list(l2.values()))
    worst = sorted(l2.items(), key=lambda kv: -kv[1])[:TOPN]
    print(f"  [{tag}] rel-L2 per tensor: median {np.median(v):.3e}  p90 {np.percentile(v, 90):.3e}  max {v.max():.3e}  "
          f"(> 2e-2: {(v > 2e-2).sum()} of {v.size})")
    for k, e in worst:
        print(f"        {e:10.3e}  {k}")


def flat(grads):
    d = {"in.memory": grads["memory"], "in.tgt": grads["tgt"], "in.refpoints_unsigmoid": grads["refpoints_unsigmoid"]}
    d.update(grads["params"])
    return {k: v.double().cpu() for k, v in d.items()}


for name in sys.argv[1:] or ["grad_ava_vitb_b1_l2"]:
    g = load_golden(name)
    cfg, B, W, inp = case_from_golden(g)
    seed = int(g["meta"][8])
    tz = bool(int(g["meta"][10]))
    print(f"== {name}: {cfg}, B {B}")
    _, gA, engA = run_train(cfg, B, W, inp, seed, torch.bfloat16)
    _, l2A = grad_errors(gA, g, seed, tgt_zero=tz, want_l2=True)
    summary("A bf16 path vs reference (fp32, unrounded weights)", l2A)
    fA = flat(gA)
    del engA
    Wr = round_matrices(W, cfg["layers"])
    _, gB, engB = run_train(cfg, B, Wr, inp, seed, torch.float32)
    _, l2B = grad_errors(gB, g, seed, tgt_zero=tz, want_l2=True)
    summary("B fp32 path, bf16-rounded matrices vs reference (weight quantisation alone)", l2B)
    fB = flat(gB)
    Gmed = float(np.median([float(v.abs().max()) for v in fB.values()]))
    l2C = {}
    for k in fA:
        if ZERO_BIAS.search(k) or (tz and ZERO_TGT0.match(k)) or "q_proj." in k or k.startswith("cls_norm."):
            continue
        n = float(fB[k].norm())
        if n > 0 and float(fB[k].abs().max()) > 1e-4 * Gmed:
            l2C[k] = float((fA[k] - fB[k]).norm()) / n
    summary("C bf16 path vs fp32 path on the SAME bf16-representable matrices (arithmetic/storage error alone)", l2C)
    del engB
    torch.cuda.empty_cache()
