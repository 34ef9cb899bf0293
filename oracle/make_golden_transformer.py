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


def main():
    ref = import_reference()
    import ops.modules.ms_deform_attn as mod
    mod.MSDeformAttnFunction = _TorchCore
    c = CFG
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
    t = lambda a: torch.from_numpy(a.copy())
    masks = [torch.zeros((c["B"],) + s, dtype=torch.bool) for s in c["shapes"]]
    with torch.no_grad():
        hs, cls_hs, refs = tr([t(s) for s in srcs], masks, [t(p) for p in poss], t(refpoint))
    np.savez_compressed(os.path.join(GOLD, "transformer_tiny.npz"), hs=hs.numpy(), cls_hs=cls_hs.numpy(), refs=refs.numpy())
    print("transformer_tiny", tuple(hs.shape), tuple(cls_hs.shape), tuple(refs.shape))


if __name__ == "__main__":
    main()
