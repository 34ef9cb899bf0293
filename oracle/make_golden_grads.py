"""Generates tests/golden/grad_*.npz: gradients of the UNMODIFIED reference decoder (torch autograd, fp32, CPU, eval mode so
that dropout is the identity) for the training-step parity of BASELINE.json configs[1] / SURVEY.md section 8(d) "Config 2":

    loss = sum(w_h * hs) + sum(w_c * cls_hs) + sum(w_r * refs)          (fixed random w_*, oracle/synth.make_loss_weights)

Run in the build container only (needs /root/reference):   python -m oracle.make_golden_grads [name ...]

ReLU kinks.  A ReLU whose pre-activation lies within rounding noise of zero has no usable derivative: two correct fp32
implementations can land on different sides and their gradients then differ by a whole term (observed: ONE flipped hidden
unit of cls_linear1_ out of 2.4 M moved cls_linear1_.bias by 3e-2 of its maximum).  With millions of pre-activations a few
always sit inside +-1e-6, so the generator makes every case well-posed: `clear_relu_kinks` nudges the bias of each hidden
unit that has a pre-activation within KINK_MARGIN x rms of zero and repeats until none is left.  The nudged bias vectors
are stored in the fixture (`wb.<name>`) and tests/helpers.case_from_golden applies them on top of the seeded weights.

Fixture contents (kept small): full gradients of the inputs (memory, tgt, refpoints_unsigmoid) when they are small, else a
seeded sample; for every parameter the full gradient when it has <= 4096 elements, else `synth.grad_sample_index` samples
plus its L2 norm and its sum.  Parameters the reference never uses (grad None: decoder.cls_norm.*, cls_layers.*.q_proj.*,
SURVEY.md section 8c "Gradient oracle") are recorded as zeros.
"""
import os
import re
import sys
import numpy as np
import torch

from . import synth
from .ref_import import import_reference
from .make_golden import build_reference_decoder, GOLD

# (fixture, config, B, layers override, seed, masked, tgt_zero)
GRAD_CASES = [
    ("grad_tiny", "tiny", 2, None, 0, False, True),
    ("grad_tiny_masked", "tiny", 2, None, 1, True, False),
    ("grad_small_masked", "small", 3, None, 2, True, True),
    ("grad_jhmdb_like", dict(nq=5, tprime=2, h=16, w=16, K=21, layers=1, F=2048), 1, None, 5, False, True),
    ("grad_ava_vitb_b1_l2", "ava_vitb", 1, 2, 0, False, True),      # BASELINE shape (nq 15, S 196, K 80, F 2048), 2 layers
    ("grad_ava_csn_b1_l1", "ava_csn152", 1, 1, 3, True, True),      # CSN-152 grid 16x16 = 256 keys (the largest S), masked
    ("grad_ucf_like", dict(nq=15, tprime=2, h=14, w=14, K=24, layers=1, F=2048), 1, None, 4, False, True),
    # the benchmarked depth: full 6-layer AVA22_ViT-B decoder (BASELINE configs[1] shape), 2 clips
    ("grad_ava_vitb_b2_l6", "ava_vitb", 2, None, 0, False, True),
]


# Linear modules whose output feeds a ReLU (dab_transformer.py:47 MLP, :993 linear1, :1045 cls_linear1, :1076 cls_linear1_)
RELU_FEEDERS = re.compile(r"(^|\.)(linear1|cls_linear1|cls_linear1_)$|^(query_scale|ref_point_head|ref_anchor_head)\.layers\.0$"
                          r"|^bbox_embed\.layers\.[01]$")
KINK_MARGIN = 2e-4
# 6 layers x 2 clips have 3e7 ReLU pre-activations: a 2e-4 margin cannot be cleared by bias nudges (every hidden unit has rows
# inside it).  This fixture anchors the bf16 tolerance (2e-2) and relative-L2 figures, so only pre-activations within fp32
# rounding of zero are moved.
KINK_MARGIN_BY_CASE = {"grad_ava_vitb_b2_l6": 4e-6}


def clear_relu_kinks(dec, forward, max_iter=40, margin=None):
    """Nudges ReLU-feeding biases until no pre-activation is within KINK_MARGIN * rms of zero.  Returns {bias name: vector}."""
    feeders = {n: m for n, m in dec.named_modules() if RELU_FEEDERS.search(n)}
    assert len(feeders) >= 8, sorted(feeders)
    rs = np.random.RandomState(77)
    touched = {}
    for it in range(max_iter):
        seen = {n: [] for n in feeders}
        hooks = [m.register_forward_hook(lambda mod, a, out, n=n: seen[n].append(out.detach())) for n, m in feeders.items()]
        with torch.no_grad():
            forward()
        for h in hooks:
            h.remove()
        bad_total = 0
        for n, outs in seen.items():
            if not outs:      # e.g. query_scale in a 1-layer decoder (dab_transformer.py:752: layer 0 uses scale 1)
                continue
            o = torch.cat([x.reshape(-1, x.shape[-1]) for x in outs], 0)
            thr = (margin or KINK_MARGIN) * float(o.pow(2).mean().sqrt())
            bad = (o.abs() < thr).any(0).nonzero().flatten().numpy()
            if bad.size:
                bad_total += int((o.abs() < thr).sum())
                b = feeders[n].bias.data
                b[bad] += torch.from_numpy((4 * thr * rs.choice([-1.0, 1.0], size=bad.size)).astype(np.float32))
                touched[n + ".bias"] = b
        print("   kink pass", it, "pre-activations inside the margin:", bad_total)
        if bad_total == 0:
            return {k: v.numpy().copy() for k, v in touched.items()}
    raise RuntimeError("ReLU kinks not cleared")


def run_case(ref, name, cfg, B, layers, seed, masked, tgt_zero):
    c = dict(synth.CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    if layers is not None:
        c["layers"] = layers
    W = synth.make_decoder_weights(c["K"], c["layers"], c["F"], seed=seed)
    inp = synth.make_decoder_inputs(c, B, seed=seed, masked=masked, tgt_zero=tgt_zero)
    lw = synth.make_loss_weights(c, B, seed=seed)
    dec = build_reference_decoder(ref, c, W)
    t = lambda a: torch.from_numpy(a.copy())
    nudged = clear_relu_kinks(dec, lambda: dec(t(inp["tgt"]), t(inp["memory"]), memory_key_padding_mask=t(inp["mask"]),
                                               pos=t(inp["pos"]), refpoints_unsigmoid=t(inp["refpoints_unsigmoid"]),
                                               orig_res=inp["orig_res"]),
                              margin=KINK_MARGIN_BY_CASE.get(name))
    tgt = t(inp["tgt"]).requires_grad_(True)
    memory = t(inp["memory"]).requires_grad_(True)
    ref_u = t(inp["refpoints_unsigmoid"]).requires_grad_(True)
    hs, cls_hs, refs = dec(tgt, memory, memory_key_padding_mask=t(inp["mask"]), pos=t(inp["pos"]),
                           refpoints_unsigmoid=ref_u, orig_res=inp["orig_res"])
    loss = (t(lw["w_hs"]) * hs).sum() + (t(lw["w_cls"]) * cls_hs).sum() + (t(lw["w_refs"]) * refs).sum()
    loss.backward()
    out = {"loss": np.array(loss.item(), dtype=np.float64)}
    for k, v in nudged.items():
        out["wb." + k] = v
    for nm, ten in (("memory", memory), ("tgt", tgt), ("refpoints_unsigmoid", ref_u)):
        g = ten.grad.numpy()
        if g.size <= 1 << 16:
            out["gin." + nm] = g
        else:
            idx = synth.grad_sample_index(g.size, seed)
            out["gin_s." + nm] = g.reshape(-1)[idx]
            out["gin_n." + nm] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), g.astype(np.float64).sum()])
    params = dict(dec.named_parameters())
    for nm in W:
        if nm.startswith("heads.") or ".conv_blocks.1." in nm or ".conv_blocks.2." in nm:
            continue
        p = params[nm]
        g = np.zeros(tuple(p.shape), dtype=np.float32) if p.grad is None else p.grad.numpy()
        if g.size <= 4096:
            out["g." + nm] = g
        else:
            idx = synth.grad_sample_index(g.size, seed)
            out["gs." + nm] = g.reshape(-1)[idx]
            out["gn." + nm] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), g.astype(np.float64).sum()])
    out["meta"] = np.array([B, c["nq"], c["tprime"], c["h"], c["w"], c["K"], c["layers"], c["F"], seed, int(masked),
                            int(tgt_zero)], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "loss", loss.item(), "entries", len(out))


def main():
    torch.set_num_threads(os.cpu_count())
    sel = set(sys.argv[1:])
    ref = import_reference()
    for case in GRAD_CASES:
        if not sel or any(s in case[0] for s in sel):
            run_case(ref, *case)


if __name__ == "__main__":
    main()
