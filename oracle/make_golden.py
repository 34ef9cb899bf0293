"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) on the
deterministic synthetic inputs/weights of oracle/synth.py.  Run in the build container only:

    python -m oracle.make_golden            # all fixtures
    python -m oracle.make_golden tiny       # a subset

fp32, eval mode, CPU, torch.set_num_threads(all).  The committed fixtures are what pins oracle/decoder_np.py
(tests/test_oracle_golden.py) and, transitively, the CUDA path (tests/test_decoder_gpu.py).
"""
import os
import sys
import numpy as np
import torch

from . import synth
from .ref_import import import_reference

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (fixture name, config, B, layers override, seed, masked, tgt_zero)
DECODER_CASES = [
    ("dec_tiny", "tiny", 2, None, 0, False, True),
    ("dec_tiny_masked", "tiny", 2, None, 1, True, False),
    ("dec_small_masked", "small", 3, None, 2, True, True),
    ("dec_ava_vitb_b1", "ava_vitb", 1, None, 0, False, True),      # BASELINE.json configs[0] (1a)
    ("dec_ava_csn_b1_l2", "ava_csn152", 1, 2, 3, False, True),     # 16x16 grid, 2 layers
    ("dec_ucf_like", dict(nq=15, tprime=2, h=14, w=14, K=24, layers=1, F=2048), 1, None, 4, False, True),
    ("dec_jhmdb_like", dict(nq=5, tprime=2, h=16, w=16, K=21, layers=1, F=2048), 1, None, 5, False, True),
]


def build_reference_decoder(ref, cfg, W):
    c = cfg
    tr = ref.Transformer(d_model=256, dropout=0.1, nhead=8, num_queries=c["nq"], dim_feedforward=c["F"],
                         num_encoder_layers=1, num_decoder_layers=c["layers"], num_feature_levels=4, enc_n_points=8,
                         return_intermediate_dec=True, query_dim=4, num_classes=c["K"], temp_len=16)
    dec = tr.decoder
    dec.bbox_embed = ref.MLP(256, 256, 4, 3)          # models/model.py:90,100-101
    sd = {k: torch.from_numpy(v.copy()) for k, v in W.items() if not k.startswith("heads.")}
    missing, unexpected = dec.load_state_dict(sd, strict=True)
    dec.eval()
    return dec


def run_decoder_case(ref, name, cfg, B, layers, seed, masked, tgt_zero):
    c = dict(synth.CONFIGS[cfg]) if isinstance(cfg, str) else dict(cfg)
    if layers is not None:
        c["layers"] = layers
    W = synth.make_decoder_weights(c["K"], c["layers"], c["F"], seed=seed)
    inp = synth.make_decoder_inputs(c, B, seed=seed, masked=masked, tgt_zero=tgt_zero)
    dec = build_reference_decoder(ref, c, W)
    taps = {}

    def hook_loc(i):
        def f(mod, args, out):
            taps[f"l{i}.output"] = out[0].detach().numpy().copy()
            taps[f"l{i}.actor"] = out[1].detach().numpy().copy()
        return f

    def hook_cls(i):
        def f(mod, args, out):
            taps[f"l{i}.cls_output"] = out[0].detach().numpy().copy()
        return f

    small = c["h"] * c["w"] * c["nq"] * B * c["tprime"] <= 2000
    if small:
        for i, (l, cl) in enumerate(zip(dec.layers, dec.cls_layers)):
            l.register_forward_hook(hook_loc(i)); cl.register_forward_hook(hook_cls(i))
    t = lambda a: torch.from_numpy(a)
    with torch.no_grad():
        hs, cls_hs, refs = dec(t(inp["tgt"]), t(inp["memory"]), memory_key_padding_mask=t(inp["mask"]),
                               pos=t(inp["pos"]), refpoints_unsigmoid=t(inp["refpoints_unsigmoid"]),
                               orig_res=inp["orig_res"])
        # heads: models/model.py:191-221 restated with the reference's own modules/functions
        import torch.nn.functional as F
        from utils.misc import inverse_sigmoid
        logits_b = F.linear(hs, t(W["heads.class_embed_b.weight"]), t(W["heads.class_embed_b.bias"]))
        tmp = dec.bbox_embed(hs)
        tmp[..., :4] += inverse_sigmoid(refs)
        boxes = tmp.sigmoid()
        logits = cls_hs.mean(dim=-1)
    out = dict(hs=hs.numpy(), refs=refs.numpy(), pred_logits=logits.numpy(), pred_boxes=boxes.numpy(),
               pred_logits_b=logits_b.numpy())
    cls = cls_hs.numpy()
    if cls.size * 4 <= 1 << 20:
        out["cls_hs"] = cls
    else:  # keep fixtures small: strided sample + per-(layer,b,n,k) first/second moments
        out["cls_hs_sub"] = np.ascontiguousarray(cls[:, :, ::4, ::7, ::5])
        out["cls_hs_sq"] = (cls.astype(np.float64) ** 2).mean(-1).astype(np.float32)
    out.update({k: v for k, v in taps.items()})
    out["meta"] = np.array([B, c["nq"], c["tprime"], c["h"], c["w"], c["K"], c["layers"], c["F"], seed, int(masked),
                            int(tgt_zero)], dtype=np.int64)
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, {k: v.shape for k, v in out.items()})


def run_posenc():
    sys.path.insert(0, os.environ.get("CQVAD_REF", "/root/reference"))
    from models.position_encoding import PositionEmbeddingSine_3D
    from utils.misc import NestedTensor
    pe = PositionEmbeddingSine_3D(256, normalize=True)
    rs = np.random.RandomState(7)
    mask = np.zeros((2, 3, 5, 6), dtype=bool)
    mask[1, :, 4:, :] = True
    mask[1, :, :, 4:] = True
    x = torch.zeros(2, 256, 3, 5, 6)
    out = pe(NestedTensor(x, torch.from_numpy(mask))).numpy()
    mask2 = np.zeros((1, 8, 14, 14), dtype=bool)
    out2 = pe(NestedTensor(torch.zeros(1, 256, 8, 14, 14), torch.from_numpy(mask2))).numpy()
    ref = import_reference()
    r = torch.from_numpy(rs.uniform(0.02, 0.98, size=(4, 3, 4)).astype(np.float32))
    sine = ref.gen_sineembed_for_position(r).numpy()
    np.savez_compressed(os.path.join(GOLD, "posenc.npz"), mask=mask, pos=out, pos_vit14=out2[:, :, ::3, ::5, ::4],
                        ref_in=r.numpy(), sine=sine)
    print("posenc", out.shape, out2.shape, sine.shape)


def run_attention(ref):
    """models/detr/attention.py MultiheadAttention in its three call modes (module-level drop-in parity)."""
    from models.detr.attention import MultiheadAttention
    rs = np.random.RandomState(11)
    out = {}
    torch.manual_seed(0)
    # mode A, E=Ev=256 (self-attention), with key_padding_mask
    m = MultiheadAttention(256, 8, dropout=0.1, vdim=256).eval()
    wo = rs.uniform(-0.1, 0.1, (256, 256)).astype(np.float32); bo = (0.02 * rs.standard_normal(256)).astype(np.float32)
    m.out_proj.weight.data = torch.from_numpy(wo); m.out_proj.bias.data = torch.from_numpy(bo)
    q = rs.standard_normal((7, 3, 256)).astype(np.float32); k = rs.standard_normal((9, 3, 256)).astype(np.float32)
    v = rs.standard_normal((9, 3, 256)).astype(np.float32)
    kpm = np.zeros((3, 9), dtype=bool); kpm[1, 6:] = True
    with torch.no_grad():
        o, wts = m(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), key_padding_mask=torch.from_numpy(kpm))
    out.update(a_q=q, a_k=k, a_v=v, a_kpm=kpm, a_wo=wo, a_bo=bo, a_out=o.numpy(), a_w=wts.numpy())
    # mode A, E=512, Ev=256 (class cross-attention)
    m = MultiheadAttention(512, 8, dropout=0.1, vdim=256).eval()
    m.out_proj.weight.data = torch.from_numpy(wo); m.out_proj.bias.data = torch.from_numpy(bo)
    q = rs.standard_normal((5, 4, 512)).astype(np.float32); k = rs.standard_normal((12, 4, 512)).astype(np.float32)
    v = rs.standard_normal((12, 4, 256)).astype(np.float32)
    with torch.no_grad():
        o, _ = m(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v))
    out.update(c_q=q, c_k=k, c_v=v, c_out=o.numpy())
    # mode B, query_specific_key
    m = MultiheadAttention(512, 8, dropout=0.1, vdim=256, query_specific_key=True).eval()
    m.out_proj.weight.data = torch.from_numpy(wo); m.out_proj.bias.data = torch.from_numpy(bo)
    q = rs.standard_normal((3, 2, 512)).astype(np.float32); k = rs.standard_normal((3, 10, 2, 512)).astype(np.float32)
    v = rs.standard_normal((3, 10, 2, 256)).astype(np.float32)
    kpm = np.zeros((2, 10), dtype=bool); kpm[0, 7:] = True
    with torch.no_grad():
        o, _ = m(torch.from_numpy(q), torch.from_numpy(k), torch.from_numpy(v), key_padding_mask=torch.from_numpy(kpm))
    out.update(b_q=q, b_k=k, b_v=v, b_kpm=kpm, b_out=o.numpy())
    np.savez_compressed(os.path.join(GOLD, "attention.npz"), **out)
    print("attention ok")


def run_msda():
    """No CPU implementation of the 3-D op exists in the reference; anchor = torch 5-D grid_sample, the 3-D
    analogue of ops/functions/ms_deform_attn_func.py:48-68 (fp64 forward + autograd gradients)."""
    import torch.nn.functional as F
    cases = {"a": dict(N=2, shapes=[(2, 3, 4), (1, 2, 2)], M=2, D=4, Lq=5, P=2, seed=0),
             "b": dict(N=1, shapes=[(4, 6, 5), (4, 3, 3), (2, 3, 2), (2, 2, 1)], M=8, D=32, Lq=37, P=8, seed=1)}
    out = {}
    for tag, kw in cases.items():
        d = synth.make_msda_inputs(**kw)
        value = torch.from_numpy(d["value"]).double().requires_grad_(True)
        loc = torch.from_numpy(d["loc"]).double().requires_grad_(True)
        attn = torch.from_numpy(d["attn"]).double().requires_grad_(True)
        N, Len, M, D = value.shape
        _, Lq, _, L, P, _ = loc.shape
        grids = 2 * loc - 1
        samp = []
        for l, (T, H, Wd) in enumerate(d["shapes"].tolist()):
            s0 = int(d["level_start"][l])
            v_l = value[:, s0:s0 + T * H * Wd].flatten(2).transpose(1, 2).reshape(N * M, D, T, H, Wd)
            g_l = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)[:, :, :, None, :]     # [N*M, Lq, P, 1, 3] (x,y,t)
            s = F.grid_sample(v_l, g_l, mode="bilinear", padding_mode="zeros", align_corners=False)  # [N*M,D,Lq,P,1]
            samp.append(s[..., 0])
        a = attn.transpose(1, 2).reshape(N * M, 1, Lq, L * P)
        o = (torch.stack(samp, dim=-2).flatten(-2) * a).sum(-1).view(N, M * D, Lq).transpose(1, 2).contiguous()
        go = torch.from_numpy(np.random.RandomState(5).standard_normal(o.shape))
        (o * go).sum().backward()
        out.update({f"{tag}_out": o.detach().numpy(), f"{tag}_go": go.numpy(), f"{tag}_gvalue": value.grad.numpy(),
                    f"{tag}_gloc": loc.grad.numpy(), f"{tag}_gattn": attn.grad.numpy(),
                    f"{tag}_kw": np.array([kw["N"], kw["M"], kw["D"], kw["Lq"], kw["P"], kw["seed"]], dtype=np.int64),
                    f"{tag}_shapes": d["shapes"]})
    np.savez_compressed(os.path.join(GOLD, "msda.npz"), **out)
    print("msda ok")


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    sel = set(sys.argv[1:])
    ref = import_reference()
    for case in DECODER_CASES:
        if not sel or any(s in case[0] for s in sel):
            run_decoder_case(ref, *case)
    if not sel or "posenc" in sel:
        run_posenc()
    if not sel or "attention" in sel:
        run_attention(ref)
    if not sel or "msda" in sel:
        run_msda()


if __name__ == "__main__":
    main()
