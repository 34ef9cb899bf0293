"""Whole-chain backward (BASELINE configs[4] route): level flatten (+ level_embed) -> deformable encoder -> inter-stage resample
-> decoder through the drop-in `Transformer`, against autograd of the UNMODIFIED reference Transformer.forward
(tests/golden/transformer_tiny_grad.npz, oracle/make_golden_transformer.py::main_grad)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err, TOL_FP32
from oracle import synth

pytestmark = pytest.mark.gpu


def _build(c, dev, enc_layers=1):
    from class_query_vad_b200 import Transformer
    from oracle.make_golden_transformer import make_inputs
    srcs, poss, level_embed, refpoint = make_inputs(c)
    We = synth.make_encoder_layer_weights(c["F"], 4, c["P"], seed=c["seed"])
    Wd = synth.make_decoder_weights(c["K"], c["layers"], c["F"], seed=c["seed"])
    tr = Transformer(num_queries=c["nq"], num_encoder_layers=enc_layers, num_decoder_layers=c["layers"], dim_feedforward=c["F"],
                     enc_n_points=c["P"], num_classes=c["K"], temp_len=c["T"])
    sd = {"level_embed": torch.from_numpy(level_embed)}
    for l in range(enc_layers):
        sd.update({f"encoder.layers.{l}." + k: torch.from_numpy(v) for k, v in We.items()})
    sd.update({"decoder." + k: torch.from_numpy(v) for k, v in Wd.items() if not k.startswith("heads.")})
    tr.load_state_dict(sd, strict=True)
    return tr.to(dev).eval(), srcs, poss, refpoint


def test_transformer_backward_matches_reference_autograd():
    from oracle.make_golden_transformer import CFG as c, loss_weights
    g = load_golden("transformer_tiny_grad")
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    tr, srcs, poss, refpoint = _build(c, dev)
    tr.decoder.compute_dtype = torch.float32
    masks = [torch.zeros((c["B"],) + s, dtype=torch.bool, device=dev) for s in c["shapes"]]
    xs = [t(s).requires_grad_(True) for s in srcs]
    rp = t(refpoint).requires_grad_(True)
    hs, cls_hs, refs = tr(xs, masks, [t(p) for p in poss], rp)
    lw = loss_weights(c)
    loss = (t(lw["w_hs"]) * hs).sum() + (t(lw["w_cls"]) * cls_hs).sum() + (t(lw["w_refs"]) * refs).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - float(g["loss"])) < 1e-3 * max(1.0, abs(float(g["loss"])))
    got = {f"in.srcs.{l}": x.grad for l, x in enumerate(xs)}
    got["in.refpoint_embed"] = rp.grad
    unused = set(g["unused"].tolist())
    for n, p in tr.named_parameters():
        if n in unused:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
        else:
            assert p.grad is not None, f"no gradient reached {n}"
            got[n] = p.grad
    # analytically-zero gradients (key-side biases under a softmax; layer-0 q/k projections of tgt = 0) hold rounding noise in the
    # fixture: they are compared on the scale of the other gradients
    G = float(np.median([np.abs(g[k]).max() for k in g if k.startswith(("g.", "gs."))]))
    import re
    zero = re.compile(r"(sa_kcontent_proj|sa_kpos_proj|ca_kcontent_proj|ca_kpos_proj|k_proj)\.bias$|decoder\.layers\.0\.sa_(qcontent|qpos|kcontent|kpos)_proj\.")
    bad = {}
    for name, gt in got.items():
        a = gt.detach().float().cpu().numpy()
        fl = G if zero.search(name) else 0.0
        if "g." + name in g:
            ref = g["g." + name]
            err = np.abs(a - ref).max() / max(np.abs(ref).max(), fl, 1e-12)
        else:
            ref_s, (ref_norm, ref_max) = g["gs." + name], g["gn." + name]
            idx = synth.grad_sample_index(a.size, c["seed"])
            err = max(np.abs(a.reshape(-1)[idx] - ref_s).max() / max(ref_max, fl, 1e-12),
                      abs(np.sqrt((a.astype(np.float64) ** 2).sum()) - ref_norm) / max(ref_norm, fl * np.sqrt(a.size), 1e-12))
        if not err < TOL_FP32:
            bad[name] = float(err)
    assert not bad, f"gradient rel errors above {TOL_FP32}: {bad}"
    # level_embed[L-2] also reaches the decoder as `pos`, but only on the key side of softmax attentions: a per-level constant
    # shifts every key equally, so that route carries no gradient -- the row still matches the reference
    assert rel_err(tr.level_embed.grad.cpu().numpy()[2], g["g.level_embed"][2]) < TOL_FP32


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_glue_backward_kernels_match_torch_autograd(dtype):
    """cqvad_level_to_tokens_backward / cqvad_encoder_to_decoder_memory_backward against torch autograd of the reference
    formulation (flatten + transpose + level_embed; F.grid_sample) on a ragged pyramid, both grid_sample branches."""
    import torch.nn.functional as F
    from class_query_vad_b200 import flatten_levels, encoder_to_decoder_memory
    dev = torch.device("cuda:0")
    gen = torch.Generator(device="cpu").manual_seed(3)
    tol = 1e-5 if dtype == torch.float32 else 1.5e-2
    for shapes, nf in (([(4, 6, 5), (4, 3, 3), (2, 3, 4), (2, 2, 1)], 8), ([(4, 6, 6), (4, 3, 3), (4, 2, 2), (4, 1, 1)], 4)):
        B, L = 2, len(shapes)
        srcs = [torch.randn((B, 256) + s, generator=gen).to(dev).to(dtype).requires_grad_(True) for s in shapes]
        poss = [torch.randn((B, 256) + s, generator=gen).to(dev).to(dtype) for s in shapes]
        le = torch.randn((L, 256), generator=gen).to(dev).requires_grad_(True)
        src_flat, pos_flat, sh, ls = flatten_levels(srcs, poss, le)
        mem, pos0 = encoder_to_decoder_memory(src_flat + pos_flat, pos_flat, sh, ls, num_frames=nf, eff=True)
        w = torch.randn(mem.shape, generator=gen).to(dev)
        (mem.float() * w).sum().backward()
        got = [s.grad.float().clone() for s in srcs] + [le.grad.clone()]
        # reference formulation in torch (fp32 autograd): dab_transformer.py:316-321,356-365,239-294,381,391
        xs = [s.detach().float().requires_grad_(True) for s in srcs]
        le2 = le.detach().clone().requires_grad_(True)
        Tt, H, W = shapes[L - 2]
        rows = []
        for l, (x, pe) in enumerate(zip(xs, poss)):
            v = x + pe.float() + le2[l].view(1, -1, 1, 1, 1)          # (src + lvl_pos) un-flattened again
            Tl = shapes[l][0]
            if Tl == nf:
                dh, dw = torch.linspace(-1, 1, H, device=dev), torch.linspace(-1, 1, W, device=dev)
                my, mx = torch.meshgrid(dh, dw, indexing="ij")
                grid = torch.stack((my, mx), 2)[None].repeat(B * Tl, 1, 1, 1)          # the reference's (meshy, meshx) order
                o = F.grid_sample(v.permute(0, 2, 1, 3, 4).flatten(0, 1), grid, align_corners=False)
                o = o.view(B, Tl, 256, H, W).permute(0, 2, 1, 3, 4)
            else:
                dt_, dh, dw = (torch.linspace(-1, 1, n, device=dev) for n in (nf, H, W))
                mt, my, mx = torch.meshgrid(dt_, dh, dw, indexing="ij")
                grid = torch.stack((mx, my, mt), -1)[None].repeat(B, 1, 1, 1, 1)
                o = F.grid_sample(v, grid, align_corners=False)
            rows.append(o[:, :, nf // 2].flatten(2).permute(2, 0, 1))          # [(H W), B, C]
        ref_mem = torch.stack(rows, 0)
        assert rel_err(mem.detach().float().cpu().numpy(), ref_mem.detach().cpu().numpy()) < max(tol, 1e-5) * 2
        (ref_mem * w).sum().backward()
        for a, b in zip(got, [x.grad for x in xs] + [le2.grad]):
            assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < tol
