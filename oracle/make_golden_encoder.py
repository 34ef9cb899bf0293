"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/enc_*.npz: outputs of the UNMODIFIED reference
`DeformableTransformerEncoderLayer` / `DeformableTransformerEncoder.get_reference_points` (models/detr/dab_transformer.py:425-523)
and `MSDeformAttn3D` module (ops/modules/ms_deform_attn.py:122-203), fp32, CPU, eval mode (dropout = identity).

The only substitution: the module's sampling core `MSDeformAttnFunction` is CUDA-only in the reference
(ops/functions/ms_deform_attn_func.py:18-45, no CPU kernel: ops/src/cpu/ms_deform_attn_cpu.cpp:26 raises), so the module's
global `MSDeformAttnFunction` is pointed at the torch 5-D `F.grid_sample` formulation -- the 3-D analogue of the reference's own
`ms_deform_attn_core_pytorch` (ops/functions/ms_deform_attn_func.py:48-68) and the same anchor tests/golden/msda.npz uses
for the op itself.  Everything around the core (four linears, softmax, the (T, W, H) offset normaliser, residuals, LayerNorms,
FFN, reference points) is the reference's code.

Run in the build container only:   python -m oracle.make_golden_encoder
"""
import os
import sys
import numpy as np
import torch
import torch.nn.functional as F

from .ref_import import import_reference
from .make_golden import GOLD
from . import synth


class _TorchCore:
    """Stand-in for MSDeformAttnFunction: .apply(value, shapes, level_start, loc, attn, im2col_step)."""

    @staticmethod
    def apply(value, shapes, level_start, loc, attn, im2col_step):
        N, Len, M, D = value.shape
        _, Lq, _, L, P, _ = loc.shape
        grids = 2 * loc - 1
        samp = []
        for l, (T, H, Wd) in enumerate(shapes.tolist()):
            s0 = int(level_start[l])
            v_l = value[:, s0:s0 + T * H * Wd].flatten(2).transpose(1, 2).reshape(N * M, D, T, H, Wd)
            g_l = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)[:, :, :, None, :]
            samp.append(F.grid_sample(v_l, g_l, mode="bilinear", padding_mode="zeros", align_corners=False)[..., 0])
        a = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
        return (torch.stack(samp, dim=-2).flatten(-2) * a).sum(-1).view(N, M * D, Lq).transpose(1, 2).contiguous()


# (fixture, B, shapes (T,H,W) per level, d_ffn, n_points, seed, masked)
ENC_CASES = [
    ("enc_tiny", 2, [(2, 4, 3), (2, 3, 2), (1, 2, 2), (1, 1, 2)], 128, 8, 0, False),
    ("enc_small_masked", 2, [(4, 6, 5), (4, 3, 3), (2, 3, 2), (2, 2, 1)], 256, 8, 1, True),
    ("enc_mid_f2048", 1, [(4, 14, 14), (4, 7, 7), (2, 4, 4), (1, 2, 2)], 2048, 8, 2, False),   # level 0 = the decoder's 14x14 grid
]


def run_case(ref, name, B, shapes, F_, P, seed, masked):
    import ops.modules.ms_deform_attn as mod
    mod.MSDeformAttnFunction = _TorchCore
    L = len(shapes)
    W = synth.make_encoder_layer_weights(F_, L, P, seed=seed)
    inp = synth.make_encoder_inputs(B, shapes, seed=seed, masked=masked)
    layer = ref.DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, dropout=0.1, activation="relu", n_levels=L, n_heads=8,
                                                  n_points=P)
    missing, unexpected = layer.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()}, strict=True)
    layer.eval()
    t = lambda a: torch.from_numpy(a.copy())
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    refp = ref.DeformableTransformerEncoder.get_reference_points(sh, t(inp["valid_ratios"]), device="cpu")
    taps = {}
    h = layer.self_attn.register_forward_hook(lambda m, a, o: taps.__setitem__("attn_out", o.detach().numpy().copy()))
    with torch.no_grad():
        out = layer(t(inp["src"]), t(inp["pos"]), refp, sh, lsi, t(inp["mask"]) if masked else None)
    h.remove()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), out=out.numpy(), attn_out=taps["attn_out"],
                        reference_points=refp.numpy(), shapes=np.array(shapes, dtype=np.int64), level_start=lsi.numpy(),
                        meta=np.array([B, F_, P, seed, int(masked)], dtype=np.int64))
    print(name, "Len", int(sh.prod(1).sum()), "out |max|", float(out.abs().max()))


# (fixture, B, shapes, d_ffn, n_points, seed, masked): gradients of loss = sum(w * out) through the reference layer (autograd
# through the grid_sample core = the mathematical gradient the CUDA backward implements, SURVEY.md section 8a "MSDA backward")
ENC_GRAD_CASES = [
    ("enc_grad_tiny", 2, [(2, 4, 3), (2, 3, 2), (1, 2, 2), (1, 1, 2)], 128, 8, 3, False),
    ("enc_grad_small_masked", 2, [(4, 6, 5), (4, 3, 3), (2, 3, 2), (2, 2, 1)], 128, 8, 4, True),
]
KINK_MARGIN = 2e-4


def run_grad_case(ref, name, B, shapes, F_, P, seed, masked):
    import ops.modules.ms_deform_attn as mod
    mod.MSDeformAttnFunction = _TorchCore
    L = len(shapes)
    W = synth.make_encoder_layer_weights(F_, L, P, seed=seed)
    inp = synth.make_encoder_inputs(B, shapes, seed=seed, masked=masked)
    layer = ref.DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, dropout=0.1, activation="relu", n_levels=L, n_heads=8,
                                                  n_points=P)
    layer.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in W.items()}, strict=True)
    layer.eval()
    t = lambda a: torch.from_numpy(a.copy())
    sh = torch.tensor(shapes, dtype=torch.long)
    lsi = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    refp = ref.DeformableTransformerEncoder.get_reference_points(sh, t(inp["valid_ratios"]), device="cpu")
    mask = t(inp["mask"]) if masked else None
    # ReLU-kink clearing (see oracle/make_golden_grads.py): nudge linear1.bias until no pre-activation sits within the margin
    rs = np.random.RandomState(99)
    for it in range(40):
        seen = []
        hk = layer.linear1.register_forward_hook(lambda m, a, o: seen.append(o.detach()))
        with torch.no_grad():
            layer(t(inp["src"]), t(inp["pos"]), refp, sh, lsi, mask)
        hk.remove()
        o = seen[0].reshape(-1, F_)
        thr = KINK_MARGIN * float(o.pow(2).mean().sqrt())
        bad = (o.abs() < thr).any(0).nonzero().flatten().numpy()
        print("   kink pass", it, int((o.abs() < thr).sum()))
        if bad.size == 0:
            break
        layer.linear1.bias.data[bad] += torch.from_numpy((4 * thr * rs.choice([-1.0, 1.0], size=bad.size)).astype(np.float32))
    else:
        raise RuntimeError("ReLU kinks not cleared")
    src = t(inp["src"]).requires_grad_(True)
    pos = t(inp["pos"]).requires_grad_(True)
    out = layer(src, pos, refp, sh, lsi, mask)
    w_out = torch.from_numpy(np.random.RandomState(9000 + seed).standard_normal(tuple(out.shape)).astype(np.float32))
    loss = (w_out * out).sum()
    loss.backward()
    res = {"loss": np.array(loss.item(), dtype=np.float64), "out": out.detach().numpy(), "w_out": w_out.numpy(),
           "gin.src": src.grad.numpy(), "gin.pos": pos.grad.numpy(), "wb.linear1.bias": layer.linear1.bias.detach().numpy().copy(),
           "reference_points": refp.numpy(), "shapes": np.array(shapes, dtype=np.int64), "level_start": lsi.numpy(),
           "meta": np.array([B, F_, P, seed, int(masked)], dtype=np.int64)}
    for k, p_ in layer.named_parameters():
        res["g." + k] = p_.grad.numpy()
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **res)
    print(name, "loss", loss.item())


def main():
    ref = import_reference()
    for case in ENC_GRAD_CASES:
        if not sys.argv[1:] or any(a in case[0] for a in sys.argv[1:]):
            run_grad_case(ref, *case)
    sel = set(sys.argv[1:])
    for case in ENC_CASES:
        if not sel or any(s in case[0] for s in sel):
            run_case(ref, *case)


if __name__ == "__main__":
    main()
