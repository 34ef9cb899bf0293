"""TEST INFRASTRUCTURE ONLY.  tests/golden/transformer_tiny.npz: the UNMODIFIED reference `Transformer.forward`
(models/detr/dab_transformer.py:296-397: flatten + level_embed, deformable encoder, un-flatten, make_interpolated_features,
key-frame slice, decoder) on a tiny configuration, fp32, CPU, eval -- the integration fixture for the composition
encoder layer -> cqvad_encoder_to_decoder_memory -> decoder.  Only the CUDA-only sampling core is replaced (torch grid_sample,
see oracle/make_golden_encoder.py).  Run in the build container only:   python -m oracle.make_golden_transformer"""
import os
import numpy as np
import torch

from . import synth
from .ref_import import import_reference
from .make_golden import GOLD
from .make_golden_encoder import _TorchCore

CFG = dict(B=2, T=2, shapes=[(2, 6, 6), (2, 3, 3), (2, 4, 5), (2, 2, 2)], nq=3, K=5, layers=2, F=128, P=8, seed=11)


def make_inputs(cfg):
    rs = np.random.RandomState(10000 + cfg["seed"])
    srcs = [rs.standard_normal((cfg["B"], 256, t, h, w)).astype(np.float32) for (t, h, w) in cfg["shapes"]]
    poss = [(0.5 * rs.standard_normal((cfg["B"], 256, t, h, w))).astype(np.float32) for (t, h, w) in cfg["shapes"]]
    level_embed = rs.standard_normal((len(cfg["shapes"]), 256)).astype(np.float32)
    refpoint = rs.standard_normal((cfg["nq"], 1, 4)).astype(np.float32)
    return srcs, poss, level_embed, refpoint


def build_reference_transformer(ref, c):
    import ops.modules.ms_deform_attn as mod
    mod.MSDeformAttnFunction = _TorchCore
    tr = ref.Transformer(d_model=256, dropout=0.1, nhead=8, num_queries=c["nq"], dim_feedforward=c["F"], num_encoder_layers=1,
                         num_decoder_layers=c["layers"], num_feature_levels=4, enc_n_points=c["P"], return_intermediate_dec=True,
                         query_dim=4, num_classes=c["K"], temp_len=c["T"])
    tr.decoder.bbox_embed = ref.MLP(256, 256, 4, 3)       # models/model.py:90,100-101
    tr.eff = True                                         # set by the model builder (key-frame decoding)
    srcs, poss, level_embed, refpoint = make_inputs(c)
    We = synth.make_encoder_layer_weights(c["F"], 4, c["P"], seed=c["seed"])
    Wd = synth.make_decoder_weights(c["K"], c["layers"], c["F"], seed=c["seed"])
    sd = {"level_embed": torch.from_numpy(level_embed)}
    sd.update({"encoder.layers.0." + k: torch.from_numpy(v.copy()) for k, v in We.items()})
    sd.update({"decoder." + k: torch.from_numpy(v.copy()) for k, v in Wd.items() if not k.startswith("heads.")})
    tr.load_state_dict(sd, strict=True)
    tr.eval()
    return tr, srcs, poss, refpoint


def loss_weights(c):
    """Weights of the synthetic loss sum(w_hs*hs) + sum(w_cls*cls_hs) + sum(w_refs*refs)."""
    return synth.make_loss_weights(dict(layers=c["layers"], tprime=1, nq=c["nq"], K=c["K"]), c["B"], seed=c["seed"])


FULL_BELOW = 4096


def main_grad(ref):
    """transformer_tiny_grad.npz: autograd of the reference Transformer.forward (eval: dropout = identity) with respect to the
    pyramid levels, level_embed, refpoint_embed and every encoder / decoder parameter -- the fixture of the whole-chain backward
    (level flatten -> encoder -> resample -> decoder).  Small gradients in full (`g.`), large ones as seeded samples + norm."""
    c = CFG
    tr, srcs, poss, refpoint = build_reference_transformer(ref, c)
    t = lambda a: torch.from_numpy(a.copy())
    masks = [torch.zeros((c["B"],) + s, dtype=torch.bool) for s in c["shapes"]]
    xs = [t(s).requires_grad_(True) for s in srcs]
    rp = t(refpoint).requires_grad_(True)
    hs, cls_hs, refs = tr(xs, masks, [t(p) for p in poss], rp)
    lw = loss_weights(c)
    loss = (t(lw["w_hs"]) * hs).sum() + (t(lw["w_cls"]) * cls_hs).sum() + (t(lw["w_refs"]) * refs).sum()
    loss.backward()
    out = {"loss": np.float64(loss.item())}

    def put(name, g):
        g = g.detach().numpy()
        if g.size <= FULL_BELOW:
            out["g." + name] = g
        else:
            out["gs." + name] = g.reshape(-1)[synth.grad_sample_index(g.size, c["seed"])]
            out["gn." + name] = np.array([np.sqrt((g.astype(np.float64) ** 2).sum()), np.abs(g).max()])
    for l, x in enumerate(xs):
        put(f"in.srcs.{l}", x.grad)
    put("in.refpoint_embed", rp.grad)
    unused = []
    for n, p_ in tr.named_parameters():
        if p_.grad is None:
            unused.append(n)
            continue
        put(n, p_.grad)
    out["unused"] = np.array(unused)
    np.savez_compressed(os.path.join(GOLD, "transformer_tiny_grad.npz"), **out)
    print("transformer_tiny_grad: loss", float(loss), len(out), "entries; no-grad parameters:", unused)


def main():
    ref = import_reference()
    c = CFG
    tr, srcs, poss, refpoint = build_reference_transformer(ref, c)
    t = lambda a: torch.from_numpy(a.copy())
    masks = [torch.zeros((c["B"],) + s, dtype=torch.bool) for s in c["shapes"]]
    with torch.no_grad():
        hs, cls_hs, refs = tr([t(s) for s in srcs], masks, [t(p) for p in poss], t(refpoint))
    np.savez_compressed(os.path.join(GOLD, "transformer_tiny.npz"), hs=hs.numpy(), cls_hs=cls_hs.numpy(), refs=refs.numpy())
    print("transformer_tiny", tuple(hs.shape), tuple(cls_hs.shape), tuple(refs.shape))
    main_grad(ref)


if __name__ == "__main__":
    main()
