"""Dropout of the native training path (reference: nn.Dropout on every residual branch / FFN hidden, dab_transformer.py:499-519,
937,991,995,1043,1062,1076).  The reference's RNG stream cannot be matched, so: (1) the mask generator is tested statistically
(keep rate, scale, independence of sites / seeds, determinism), (2) p = 0 is bit-identical to the parity path, (3) with a FIXED
seed the network is a deterministic differentiable function whose native backward must equal its directional derivative
(central differences in fp32) -- which fails by O(p) if the backward's masks differed from the forward's."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import synth

pytestmark = pytest.mark.gpu


def _dropout(x, p, seed, site, res=None):
    from class_query_vad_b200 import _lib
    out = torch.empty_like(x)
    _lib.check(_lib.lib().cqvad_dropout(_lib.dtype_id(x.dtype), _lib.ptr(x), _lib.ptr(res), _lib.ptr(out), x.numel(), float(p), int(seed),
                                        int(site), _lib.stream_ptr()))
    return out


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("p", [0.1, 0.5])
def test_dropout_kernel_statistics(dtype, p):
    dev = torch.device("cuda:0")
    n = 1 << 22
    x = torch.ones(n, device=dev, dtype=dtype)
    y = _dropout(x, p, seed=7, site=3).float()
    keep = y != 0
    rate = float(keep.float().mean())
    assert abs(rate - (1 - p)) < 5 * np.sqrt(p * (1 - p) / n) + 1e-5                       # keep probability (p quantised to 1/65536)
    scale = y[keep]
    assert float((scale - scale[0]).abs().max()) == 0 and abs(float(scale[0]) - 1 / (1 - p)) < (1e-2 if dtype == torch.bfloat16 else 1e-4)
    assert abs(float(y.mean()) - 1.0) < 5 * np.sqrt(p / (1 - p) / n) + 1e-2               # E[dropout(x)] = x
    assert torch.equal(y, _dropout(x, p, seed=7, site=3).float())                          # deterministic in (seed, site)
    for other in (_dropout(x, p, seed=8, site=3), _dropout(x, p, seed=7, site=4)):         # ... and independent across seeds / sites
        agree = float(((other.float() != 0) == keep).float().mean())
        assert abs(agree - (p * p + (1 - p) * (1 - p))) < 5e-3
    k8 = keep.view(-1, 8).float()                                                          # no structure inside the 8-element Philox groups
    assert float((k8.mean(0) - (1 - p)).abs().max()) < 5e-3
    c = np.corrcoef(k8[:, 0].cpu().numpy(), k8[:, 1].cpu().numpy())[0, 1]
    assert abs(c) < 5e-3
    res = torch.full_like(x, 2.0)
    assert torch.equal(_dropout(x, p, 7, 3, res=res).float(), y + 2.0)                     # fused residual add


def _decoder_case():
    cfg = dict(synth.CONFIGS["small"])
    B, seed = 2, 3
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=seed)
    inp = synth.make_decoder_inputs(cfg, B, seed=seed, masked=True)
    lw = synth.make_loss_weights(cfg, B, seed=1)
    return cfg, W, inp, lw


def test_decoder_dropout_zero_is_the_parity_path_and_seeds_matter():
    from class_query_vad_b200 import DecoderEngine
    cfg, W, inp, lw = _decoder_case()
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    eng = DecoderEngine(W, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=torch.float32, device=dev)
    run = lambda **kw: {k: v.clone() for k, v in eng.forward_train(t(inp["tgt"]), t(inp["memory"]), t(inp["mask"]), t(inp["pos"]),
                                                                 t(inp["refpoints_unsigmoid"]), inp["orig_res"], **kw).items()
                        if torch.is_tensor(v)}
    base, zero = run(), run(dropout_p=0.0, seed=99)
    a, a2, b = run(dropout_p=0.1, seed=5), run(dropout_p=0.1, seed=5), run(dropout_p=0.1, seed=6)
    for k in ("hs", "cls_hs", "refs"):
        assert torch.equal(base[k], zero[k]), k
        assert torch.equal(a[k], a2[k]), k
        assert float((a[k] - b[k]).abs().max()) > 1e-4 and float((a[k] - base[k]).abs().max()) > 1e-4, k


def _directional_check(f_and_grad, x, rel_eps=2e-3, tol=2e-2):
    """f_and_grad(x) -> (scalar float64, grad wrt x).  Central difference along a random direction vs <grad, d>."""
    gen = torch.Generator(device="cpu").manual_seed(0)
    d = torch.randn(x.shape, generator=gen).to(x.device)
    eps = rel_eps * float(x.abs().mean())
    f0, g = f_and_grad(x)
    fp, _ = f_and_grad(x + eps * d)
    fm, _ = f_and_grad(x - eps * d)
    num = (fp - fm) / (2 * eps)
    ana = float((g.double() * d.double()).sum())
    assert abs(num - ana) <= tol * max(abs(ana), abs(num)), (num, ana)
    return num, ana


def test_decoder_dropout_forward_and_backward_match_autograd_with_the_same_masks():
    """Exact parity WITH dropout: the masks the CUDA path draws (read back through cqvad_dropout with the same seed / site / layout)
    are injected at the nine nn.Dropout sites of the torch restatement of the reference decoder (oracle/decoder_torch.py, itself
    pinned to the reference fixtures); outputs and every gradient then agree at the fp32 tolerance.  Detached reference points
    and the detached actor feature make the decoder's gradient differ from the directional derivative of its forward, hence an
    autograd oracle here rather than finite differences."""
    from class_query_vad_b200 import DecoderEngine
    from oracle import decoder_torch
    from helpers import rel_err, TOL_FP32
    cfg, W, inp, lw = _decoder_case()
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    p, seed = 0.1, 11
    nq, K, F_, C = cfg["nq"], cfg["K"], cfg["F"], 256
    BT = inp["tgt"].shape[1]
    N = nq * BT

    def drop(x, layer, k):
        site = 0x1000 + 16 * layer + k
        if k == 6:                                   # class self-attention: CUDA rows are [N, K] (layer 0: the K shared rows)
            n = K * C if layer == 0 else N * K * C
            m = _dropout(torch.ones(n, device=dev), p, seed, site).cpu()
            m = m.view(K, 1, C) if layer == 0 else m.view(N, K, C).permute(1, 0, 2)
        else:
            m = _dropout(torch.ones(x.numel(), device=dev), p, seed, site).cpu().view(x.shape)
        return x * m
    loss_ref, g_ref, gmem_ref, gtgt_ref, gref_ref = decoder_torch.train_step(W, inp, lw, cfg["layers"], drop=drop)
    eng = DecoderEngine(W, nq=nq, K=K, layers=cfg["layers"], F=F_, dtype=torch.float32, device=dev)
    out = eng.forward_train(t(inp["tgt"]), t(inp["memory"]), t(inp["mask"]), t(inp["pos"]), t(inp["refpoints_unsigmoid"]), inp["orig_res"],
                            dropout_p=p, seed=seed)
    loss = float((t(lw["w_hs"]).double() * out["hs"].double()).sum() + (t(lw["w_cls"]).double() * out["cls_hs"].double()).sum() +
                 (t(lw["w_refs"]).double() * out["refs"].double()).sum())
    g = eng.backward(t(lw["w_hs"]), t(lw["w_cls"]), t(lw["w_refs"]))
    torch.cuda.synchronize()
    assert abs(loss - loss_ref) < 1e-3 * max(1.0, abs(loss_ref))
    assert rel_err(g["memory"].cpu().numpy(), gmem_ref) < TOL_FP32
    assert rel_err(g["tgt"].cpu().numpy(), gtgt_ref) < TOL_FP32
    assert rel_err(g["refpoints_unsigmoid"].cpu().numpy(), gref_ref) < TOL_FP32
    G = float(np.median([np.abs(v).max() for v in g_ref.values()]))
    bad = {}
    for name, got in g["params"].items():
        ref = g_ref[name]
        e = float(np.abs(got.cpu().numpy() - ref).max() / max(np.abs(ref).max(), 1e-3 * G))
        if not e < TOL_FP32:
            bad[name] = e
    assert not bad, bad
    # ... and the masks matter: without them the same oracle is far away
    loss0, _, gmem0, _, _ = decoder_torch.train_step(W, inp, lw, cfg["layers"])
    assert rel_err(g["memory"].cpu().numpy(), gmem0) > 10 * TOL_FP32


def test_encoder_layer_dropout_matches_autograd_with_the_same_masks():
    """The native encoder layer in train() mode against torch autograd of the reference formulation
    (DeformableTransformerEncoderLayer.forward, dab_transformer.py:499-523, and MSDeformAttn3D.forward, ops/modules/ms_deform_attn.py:
    167-203) with the masks of the CUDA path injected at dropout1 / dropout2 / dropout3; the sampling core of the formulation is
    this library's MSDeformAttnFunction (itself pinned to the reference kernel, tests/test_msda_ref_gpu.py).  fp32, 1e-3."""
    import torch.nn.functional as F
    from class_query_vad_b200 import DeformableTransformerEncoderLayer, DeformableTransformerEncoder, MSDeformAttnFunction
    from helpers import rel_err, TOL_FP32
    dev = torch.device("cuda:0")
    shapes = [(2, 6, 6), (2, 3, 3), (2, 4, 5), (2, 2, 2)]
    F_, P, B, M, L = 128, 8, 2, 8, 4
    We = synth.make_encoder_layer_weights(F_, 4, P, seed=2)
    inp = synth.make_encoder_inputs(B, shapes, seed=2)
    layer = DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, n_levels=4, n_heads=8, n_points=P)
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in We.items()}, strict=True)
    layer = layer.to(dev).train()
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    refp = DeformableTransformerEncoder.get_reference_points(sh, torch.ones((B, 4, 3), device=dev), dev)
    pos = torch.from_numpy(inp["pos"]).to(dev)
    src0 = torch.from_numpy(inp["src"]).to(dev)
    gen = torch.Generator(device="cpu").manual_seed(5)
    wout = torch.randn(inp["src"].shape, generator=gen).to(dev)
    p = layer.dropout1.p
    assert p > 0

    layer._train_calls = 0
    s = src0.clone().requires_grad_(True)
    out = layer(s, pos, refp, sh, ls, None)
    (out * wout).sum().backward()
    seed = (layer.dropout_seed << 24) + 1
    got = {n: q.grad.clone() for n, q in layer.named_parameters()}

    W = {n: q.detach().clone().requires_grad_(True) for n, q in layer.named_parameters()}
    lin = lambda x, n: F.linear(x, W[n + ".weight"], W[n + ".bias"])
    mask = lambda x, site: x * _dropout(torch.ones(x.numel(), device=dev), p, seed, site).view(x.shape)
    s2 = src0.clone().requires_grad_(True)
    q = s2 + pos                                                                                   # with_pos_embed :496-497
    Len = s2.shape[1]
    value = lin(s2, "self_attn.value_proj").view(B, Len, M, 32)
    off = lin(q, "self_attn.sampling_offsets").view(B, Len, M, L, P, 3)
    aw = F.softmax(lin(q, "self_attn.attention_weights").view(B, Len, M, L * P), -1).view(B, Len, M, L, P)
    norm = torch.stack([sh[..., 0], sh[..., 2], sh[..., 1]], -1).float()                           # ms_deform_attn.py:190
    loc = refp[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
    samp = MSDeformAttnFunction.apply(value.contiguous(), sh, ls, loc.contiguous(), aw.contiguous(), 64)
    x = F.layer_norm(s2 + mask(lin(samp, "self_attn.output_proj"), 1), (256,), W["norm1.weight"], W["norm1.bias"])       # :507-509
    x = F.layer_norm(x + mask(lin(mask(F.relu(lin(x, "linear1")), 2), "linear2"), 3), (256,), W["norm2.weight"], W["norm2.bias"])   # :499-503
    (x * wout).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), x.detach().cpu().numpy()) < TOL_FP32
    assert rel_err(s.grad.cpu().numpy(), s2.grad.cpu().numpy()) < TOL_FP32
    for n in got:
        assert rel_err(got[n].cpu().numpy(), W[n].grad.cpu().numpy()) < TOL_FP32, n
    layer.eval()
    ev = layer(src0.clone().requires_grad_(True), pos, refp, sh, ls, None)
    assert float((ev - out).abs().max()) > 1e-3          # train() mode really drops
