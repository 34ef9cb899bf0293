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


def global_l2(fa, fb, tz):
    """relative L2 of the whole parameter gradient (all tensors concatenated) and its cosine."""
    num = den = dot = na = 0.0
    for k in fb:
        if k.startswith("in.") or k not in fa:
            continue
        a, b = fa[k], fb[k]
        num += float((a - b).pow(2).sum()); den += float(b.pow(2).sum()); dot += float((a * b).sum()); na += float(a.pow(2).sum())
    return (num / den) ** 0.5, dot / (na * den) ** 0.5


def reference_bf16_errors(g, cfg, B, W, inp, seed, tz):
    from oracle import ref_import, synth
    from oracle.make_golden import build_reference_decoder
    ref = ref_import.import_reference()
    dev = torch.device("cuda:0")
    lw = synth.make_loss_weights(cfg, B, seed=seed)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {}
    for tag in ("autocast(bf16) over the fp32 reference module", "reference module .bfloat16() (weights + activations bf16)"):
      try:
            dec = build_reference_decoder(ref, cfg, W).to(dev)
            pure = tag.startswith("reference module")
            if pure:
                dec = dec.bfloat16()
            cast = (lambda x: x.bfloat16()) if pure else (lambda x: x)
            tgt = cast(t(inp["tgt"])).requires_grad_(True)
            mem = cast(t(inp["memory"])).requires_grad_(True)
            refu = cast(t(inp["refpoints_unsigmoid"])).requires_grad_(True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=not pure):
                hs, cls_hs, refs = dec(tgt, mem, memory_key_padding_mask=t(inp["mask"]), pos=cast(t(inp["pos"])),
                                       refpoints_unsigmoid=refu, orig_res=inp["orig_res"])
            loss = (t(lw["w_hs"]) * hs.float()).sum() + (t(lw["w_cls"]) * cls_hs.float()).sum() + (t(lw["w_refs"]) * refs.float()).sum()
            loss.backward()
            params = {k: (torch.zeros_like(p_) if p_.grad is None else p_.grad).float() for k, p_ in dec.named_parameters()}
            grads = dict(memory=mem.grad.float(), tgt=tgt.grad.float(), refpoints_unsigmoid=refu.grad.float(), params=params)
            _, l2 = grad_errors(grads, g, seed, tgt_zero=tz, want_l2=True)
            out[tag] = l2
            del dec
            torch.cuda.empty_cache()
      except Exception as e:
        print(f"  [D {tag}] does not run: {str(e)[:160]}")
    return out


def flat(grads):
    d = {"in.memory": grads["memory"], "in.tgt": grads["tgt"], "in.refpoints_unsigmoid": grads["refpoints_unsigmoid"]}
    d.update(grads["params"])
    return {k: v.double().cpu() for k, v in d.items()}


RECORD = {}          # DUMP_JSON=path: per case {ours, weight_quantisation_alone, reference_autocast_bf16}: {median, p90, max}
stats = lambda l2: {"median": float(np.median(list(l2.values()))), "p90": float(np.percentile(list(l2.values()), 90)),
                    "max": float(max(l2.values())), "tensors": len(l2)}

for name in [a for a in sys.argv[1:] if not a.startswith("-")] or ["grad_ava_vitb_b1_l2"]:
    g = load_golden(name)
    cfg, B, W, inp = case_from_golden(g)
    seed = int(g["meta"][8])
    tz = bool(int(g["meta"][10]))
    print(f"== {name}: {cfg}, B {B}")
    _, gA, engA = run_train(cfg, B, W, inp, seed, torch.bfloat16)
    _, l2A = grad_errors(gA, g, seed, tgt_zero=tz, want_l2=True)
    summary("A bf16 path vs reference (fp32, unrounded weights)", l2A)
    RECORD[name] = {"ours_bf16": stats(l2A)}
    fA = flat(gA)
    del engA
    Wr = round_matrices(W, cfg["layers"])
    _, gB, engB = run_train(cfg, B, Wr, inp, seed, torch.float32)
    _, l2B = grad_errors(gB, g, seed, tgt_zero=tz, want_l2=True)
    summary("B fp32 path, bf16-rounded matrices vs reference (weight quantisation alone)", l2B)
    RECORD[name]["weight_quantisation_alone"] = stats(l2B)
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
    gl, cs = global_l2(fA, fB, tz)
    print(f"  whole parameter gradient, bf16 path vs fp32 path on the same matrices: rel-L2 {gl:.3e}, cosine {cs:.6f}")
    for k in ("in.memory", "in.tgt", "in.refpoints_unsigmoid"):
        print(f"  {k}: A {l2A.get(k, float('nan')):.3e}  B {l2B.get(k, float('nan')):.3e}  C {l2C.get(k, float('nan')):.3e}")
    cls = {"conv (GELU ConvBlock)": "conv_blocks", "ReLU FFN linear1": "linear1", "loc chain sa_/ca_": ".sa_", "norm": "norm"}
    for tag, pat in cls.items():
        va = [v for k, v in l2A.items() if pat in k]; vb = [v for k, v in l2B.items() if pat in k]; vc = [v for k, v in l2C.items() if pat in k]
        if va:
            print(f"  class {tag:28s} median A {np.median(va):.3e}  B {np.median(vb):.3e}  C {np.median(vc):.3e}  (n={len(va)})")
    del engB
    torch.cuda.empty_cache()
    # (D) the UNMODIFIED reference itself evaluated in bf16 on this GPU: (D1) torch autocast(bf16) around the fp32 module,
    # (D2) the module cast to bf16 (weights and activations bf16) -- same fixture, same metric
    try:
        for tag, l2D in reference_bf16_errors(g, cfg, B, W, inp, seed, tz).items():
            summary(f"D {tag} vs reference fp32 fixture", l2D)
            if tag.startswith("autocast"):
                RECORD[name]["reference_autocast_bf16"] = stats(l2D)
    except Exception as e:
        print("  [D] reference bf16 evaluation failed:", str(e)[:200])

if os.environ.get("DUMP_JSON"):
    import json
    RECORD["_about"] = ("per-tensor relative-L2 gradient error against the fp32 reference-autograd fixtures, measured on B200 by tools/diag_bf16_budget.py: "
                        "ours_bf16 = the tcgen05 bf16 path; weight_quantisation_alone = the fp32 path with the matrices rounded to bf16; "
                        "reference_autocast_bf16 = the UNMODIFIED reference decoder (baseline/_ref) under torch.autocast(bf16) on the same GPU")
    json.dump(RECORD, open(os.environ["DUMP_JSON"], "w"), indent=1)
