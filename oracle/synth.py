"""TEST INFRASTRUCTURE -- deterministic synthetic weights and inputs (numpy only).

Shared by oracle/make_golden.py (loads them into the *reference* decoder), by tests/ and by bench.py, so the
reference run that produced tests/golden/*.npz and the CUDA path see bit-identical fp32 inputs without the
(large) weights having to be committed.  Parameter names/shapes follow the reference state_dict
(models/detr/dab_transformer.py:671-720, 854-903, 999-1038; models/model.py:87-97).

`np.random.RandomState` is used on purpose: its stream is frozen by numpy's compatibility policy.
"""
from collections import OrderedDict
import numpy as np

D_MODEL = 256
NHEAD = 8

# name -> (nq, T', h, w, K, layers, F)   (SURVEY.md shape glossary; configuration/*.yaml)
CONFIGS = {
    "ava_vitb": dict(nq=15, tprime=1, h=14, w=14, K=80, layers=6, F=2048),
    "ava_csn152": dict(nq=15, tprime=1, h=16, w=16, K=80, layers=6, F=2048),
    "ucf_vitb": dict(nq=15, tprime=32, h=14, w=14, K=24, layers=3, F=2048),
    "jhmdb_vitb": dict(nq=5, tprime=40, h=16, w=16, K=21, layers=3, F=2048),
    # tiny cases for fast parity tests (ragged extents on purpose)
    "tiny": dict(nq=3, tprime=1, h=3, w=5, K=5, layers=2, F=128),
    "small": dict(nq=4, tprime=1, h=6, w=7, K=11, layers=3, F=256),
}


def decoder_param_spec(K, layers, F, C=D_MODEL):
    """Ordered (name, shape) list of the reference TransformerDecoder state_dict (+ DETR heads under 'heads.')."""
    spec = []

    def lin(name, o, i):
        spec.append((name + ".weight", (o, i)))
        spec.append((name + ".bias", (o,)))

    def ln(name):
        spec.append((name + ".weight", (C,)))
        spec.append((name + ".bias", (C,)))

    for i in range(layers):
        p = f"layers.{i}."
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj"):
            lin(p + n, C, C)
        lin(p + "self_attn.out_proj", C, C)
        ln(p + "norm1")
        lin(p + "lvl_w_embed", 4, C)
        lin(p + "ca_qcontent_proj", C, C)
        if i == 0:  # ca_qpos_proj is None for layers >= 1 (dab_transformer.py:711-713)
            lin(p + "ca_qpos_proj", C, C)
        for n in ("ca_kcontent_proj", "ca_kpos_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            lin(p + n, C, C)
        lin(p + "cross_attn.out_proj", C, C)
        lin(p + "linear1", F, C)
        lin(p + "linear2", C, F)
        ln(p + "norm2"); ln(p + "norm3"); ln(p + "norm_")
    for i in range(layers):
        p = f"cls_layers.{i}."
        lin(p + "cls_linear1", F, C)
        lin(p + "cls_linear2", C, F)
        ln(p + "cls_norm"); ln(p + "conv_norm")
        # conv_blocks.{0,1,2} alias ONE ConvBlock (dab_transformer.py:1017-1018); generated once as conv_blocks.0
        spec.append((p + "conv_blocks.0.conv1.weight", (C, C, 3, 3)))
        spec.append((p + "conv_blocks.0.conv1.bias", (C,)))
        ln(p + "conv_blocks.0.norm")
        lin(p + "conv_blocks.0.conv2", 4 * C, C)
        lin(p + "conv_blocks.0.conv3", C, 4 * C)
        lin(p + "self_attn.out_proj", C, C)
        ln(p + "norm1")
        lin(p + "q_proj", C, C)  # unused parameter (dab_transformer.py:1026)
        spec.append((p + "k_proj.weight", (C, C, 1, 1))); spec.append((p + "k_proj.bias", (C,)))
        spec.append((p + "v_proj.weight", (C, C, 1, 1))); spec.append((p + "v_proj.bias", (C,)))
        lin(p + "cls_qpos_sine_proj", C, C)
        lin(p + "cross_attn.out_proj", C, C)
        lin(p + "cls_linear1_", F, C)
        lin(p + "cls_linear2_", C, F)
        ln(p + "cls_norm_")
    ln("norm"); ln("cls_norm"); ln("cls_norm2")
    lin("query_scale.layers.0", C, C); lin("query_scale.layers.1", C, C)
    lin("ref_point_head.layers.0", C, 2 * C); lin("ref_point_head.layers.1", C, C)
    lin("ref_anchor_head.layers.0", C, C); lin("ref_anchor_head.layers.1", 2, C)
    spec.append(("class_queries.weight", (K, C)))
    lin("bbox_embed.layers.0", C, C); lin("bbox_embed.layers.1", C, C); lin("bbox_embed.layers.2", 4, C)
    lin("heads.class_embed_b", 3, C)
    return spec


def make_decoder_weights(K, layers, F, seed=0):
    """fp32 numpy weights.  Xavier-uniform-like matrices, small non-zero biases (the reference zero-inits
    bbox_embed.layers[-1]; perturbed here so that parity is not vacuous -- SURVEY.md section 8c)."""
    rs = np.random.RandomState(1000 + seed)
    out = OrderedDict()
    for name, shape in decoder_param_spec(K, layers, F):
        leaf = name.rsplit(".", 1)[1]
        is_norm = ("norm" in name.rsplit(".", 2)[-2]) and len(shape) == 1
        if is_norm and leaf == "weight":
            v = 1.0 + 0.1 * rs.standard_normal(shape)
        elif is_norm and leaf == "bias":
            v = 0.05 * rs.standard_normal(shape)
        elif name == "class_queries.weight":
            v = rs.standard_normal(shape)  # nn.Embedding init is N(0,1); xavier is applied after -> use modest scale
            v *= 0.5
        elif leaf == "bias":
            v = 0.02 * rs.standard_normal(shape)
        else:
            fan_out = shape[0] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[0]
            fan_in = shape[1] * int(np.prod(shape[2:])) if len(shape) > 2 else shape[1]
            a = np.sqrt(6.0 / (fan_in + fan_out))
            v = rs.uniform(-a, a, size=shape)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    for i in range(layers):  # aliases
        p = f"cls_layers.{i}.conv_blocks."
        for j in (1, 2):
            for s in ("conv1.weight", "conv1.bias", "norm.weight", "norm.bias", "conv2.weight", "conv2.bias",
                      "conv3.weight", "conv3.bias"):
                out[p + f"{j}." + s] = out[p + "0." + s]
    return out


def posenc3d_key_frame(BT, h, w, t_frames=1, frame=0):
    """pos[0] of the decoder: PositionEmbeddingSine_3D (models/position_encoding.py:32-73) for an all-valid mask,
    one frame kept.  Returns [S, BT, 256] fp32."""
    from .posenc_np import position_embedding_sine_3d
    mask = np.zeros((1, t_frames, h, w), dtype=bool)
    pe = position_embedding_sine_3d(mask, num_pos_feats=D_MODEL)  # [1,256,T,H,W]
    pe = pe[0, :, frame].reshape(D_MODEL, h * w).T  # [S,256]
    return np.ascontiguousarray(np.broadcast_to(pe[:, None, :], (h * w, BT, D_MODEL)), dtype=np.float32)


def make_decoder_inputs(cfg, B, seed=0, masked=False, tgt_zero=True):
    """Synthetic decoder inputs in the layouts of TransformerDecoder.forward (dab_transformer.py:722-730, 391-396):
    tgt [nq,BT,C], memory [L,S,BT,C], pos [L,S,BT,C] (4 identical levels, dab_transformer.py:287,292),
    mask [BT,S] bool, refpoints_unsigmoid [nq,BT,4]."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    nq, tp, h, w = c["nq"], c["tprime"], c["h"], c["w"]
    BT, S = B * tp, h * w
    rs = np.random.RandomState(2000 + seed)
    memory = rs.standard_normal((4, S, BT, D_MODEL)).astype(np.float32)
    level_embed = rs.standard_normal((D_MODEL,)).astype(np.float32)
    pos0 = posenc3d_key_frame(BT, h, w) + level_embed[None, None, :]
    pos = np.ascontiguousarray(np.broadcast_to(pos0[None], (4, S, BT, D_MODEL)))
    ref = rs.standard_normal((nq, BT, 4)).astype(np.float32)
    if tgt_zero:
        tgt = np.zeros((nq, BT, D_MODEL), dtype=np.float32)
    else:
        tgt = (0.5 * rs.standard_normal((nq, BT, D_MODEL))).astype(np.float32)
    mask = np.zeros((BT, S), dtype=bool)
    if masked:  # pad the right-most columns / bottom rows of odd batch elements (never a whole row of keys)
        m = mask.reshape(BT, h, w)
        m[1::2, :, w - max(1, w // 4):] = True
        m[1::2, h - max(1, h // 5):, :] = True
    return dict(tgt=tgt, memory=memory, pos=pos, mask=mask, refpoints_unsigmoid=ref, orig_res=(h, w))


def make_msda_inputs(N, shapes, M=8, D=32, Lq=None, P=8, seed=0, spread=0.35):
    """value [N,Len,M,D], shapes [L,3] (T,H,W) int64, level_start [L] int64, loc [N,Lq,M,L,P,3] (x,y,t) in about
    [-0.2,1.2] so that out-of-range points and border corners are exercised, attn [N,Lq,M,L,P] (softmaxed)."""
    shapes = np.asarray(shapes, dtype=np.int64)
    L = shapes.shape[0]
    sizes = shapes.prod(1)
    Len = int(sizes.sum())
    lsi = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    if Lq is None:
        Lq = Len
    rs = np.random.RandomState(3000 + seed)
    value = rs.standard_normal((N, Len, M, D)).astype(np.float32)
    loc = (0.5 + spread * 2 * (rs.uniform(size=(N, Lq, M, L, P, 3)) - 0.5) * 2).astype(np.float32)
    a = rs.standard_normal((N, Lq, M, L * P)).astype(np.float32)
    a = np.exp(a - a.max(-1, keepdims=True)); a /= a.sum(-1, keepdims=True)
    attn = a.reshape(N, Lq, M, L, P).astype(np.float32)
    return dict(value=value, shapes=shapes, level_start=lsi, loc=loc, attn=attn)


def make_loss_weights(cfg, B, seed=0):
    """Fixed random weights of the synthetic training loss (Config 2):
    loss = sum(w_hs*hs) + sum(w_cls*cls_hs) + sum(w_refs*refs), shapes of TransformerDecoder.forward's outputs
    (dab_transformer.py:840-844): hs [Lr,BT,nq,C], cls_hs [Lr,BT,nq,K,C], refs [Lr,BT,nq,4]."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    BT = B * c["tprime"]
    rs = np.random.RandomState(4000 + seed)
    w_hs = rs.standard_normal((c["layers"], BT, c["nq"], D_MODEL)).astype(np.float32)
    w_cls = (rs.standard_normal((c["layers"], BT, c["nq"], c["K"], D_MODEL)) / np.sqrt(c["K"])).astype(np.float32)
    w_refs = rs.standard_normal((c["layers"], BT, c["nq"], 4)).astype(np.float32)
    return dict(w_hs=w_hs, w_cls=w_cls, w_refs=w_refs)


def grad_sample_index(size, seed=0, n=2048):
    """Seeded flat indices at which large gradients are sampled in tests/golden/grad_*.npz."""
    rs = np.random.RandomState(5000 + seed + size % 9973)
    return np.sort(rs.randint(0, size, size=min(n, size)))


# ---- deformable encoder layer (dab_transformer.py:484-523, ops/modules/ms_deform_attn.py:122-203) -------------------------
def encoder_layer_param_spec(F, L=4, P=8, M=8, C=D_MODEL):
    """Ordered (name, shape) list of the reference DeformableTransformerEncoderLayer state_dict."""
    return [("self_attn.sampling_offsets.weight", (M * L * P * 3, C)), ("self_attn.sampling_offsets.bias", (M * L * P * 3,)),
            ("self_attn.attention_weights.weight", (M * L * P, C)), ("self_attn.attention_weights.bias", (M * L * P,)),
            ("self_attn.value_proj.weight", (C, C)), ("self_attn.value_proj.bias", (C,)),
            ("self_attn.output_proj.weight", (C, C)), ("self_attn.output_proj.bias", (C,)),
            ("norm1.weight", (C,)), ("norm1.bias", (C,)),
            ("linear1.weight", (F, C)), ("linear1.bias", (F,)), ("linear2.weight", (C, F)), ("linear2.bias", (C,)),
            ("norm2.weight", (C,)), ("norm2.bias", (C,))]


def make_encoder_layer_weights(F, L=4, P=8, seed=0):
    """Deterministic non-degenerate weights: the reference init zeroes sampling_offsets.weight / attention_weights.* (parity would
    be vacuous), so every tensor is random; sampling offsets are a few voxels wide (bias = ring pattern scale, weight small)."""
    rs = np.random.RandomState(7000 + seed)
    W = {}
    for name, shape in encoder_layer_param_spec(F, L, P):
        if name.startswith("norm"):
            W[name] = (1.0 + 0.2 * rs.standard_normal(shape)).astype(np.float32) if name.endswith("weight") else \
                (0.1 * rs.standard_normal(shape)).astype(np.float32)
        elif name == "self_attn.sampling_offsets.bias":
            W[name] = (1.5 * rs.standard_normal(shape)).astype(np.float32)
        elif name == "self_attn.sampling_offsets.weight":
            W[name] = (0.05 * rs.standard_normal(shape)).astype(np.float32)
        elif name.endswith("bias"):
            W[name] = (0.1 * rs.standard_normal(shape)).astype(np.float32)
        else:
            W[name] = (rs.standard_normal(shape) / np.sqrt(shape[1])).astype(np.float32)
    return W


def make_encoder_inputs(B, shapes, seed=0, masked=False):
    """src, pos [B, Len, 256]; valid_ratios [B, L, 3] (x, y, t); mask [B, Len] bool (True = padded token)."""
    rs = np.random.RandomState(8000 + seed)
    Len = int(sum(t * h * w for (t, h, w) in shapes))
    src = rs.standard_normal((B, Len, D_MODEL)).astype(np.float32)
    pos = (0.5 * rs.standard_normal((B, Len, D_MODEL))).astype(np.float32)
    vr = np.ones((B, len(shapes), 3), dtype=np.float32)
    mask = np.zeros((B, Len), dtype=bool)
    if masked:
        vr[1::2] = (0.75 + 0.2 * rs.rand(*vr[1::2].shape)).astype(np.float32)
        mask[1::2] = rs.rand(*mask[1::2].shape) < 0.15
    return dict(src=src, pos=pos, valid_ratios=vr, mask=mask)
