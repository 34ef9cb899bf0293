"""TEST INFRASTRUCTURE -- numpy restatement of the reference class-query decoder (CPU oracle).

This file is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import it.  It is pinned against the *reference itself* (imported from
/root/reference by oracle/make_golden.py, outputs committed under tests/golden/): tests/test_oracle_golden.py.

Each function cites the reference lines it restates (paths relative to the reference root):
  mha_*                      models/detr/attention.py:190-422 (multi_head_attention_forward; projection-free)
  mlp                        models/detr/dab_transformer.py:36-48
  conv_block                 models/detr/dab_transformer.py:78-98
  decoder_layer              models/detr/dab_transformer.py:907-997   (TransformerDecoderLayer.forward)
  class_decoder_layer        models/detr/dab_transformer.py:1040-1079 (TransformerClassDecoderLayer.forward)
  decoder_forward            models/detr/dab_transformer.py:722-852   (TransformerDecoder.forward)
  detr_heads                 models/model.py:191-241
Eval-mode semantics (all dropouts are identity).  `dt` selects fp32 (the reference's arithmetic) or fp64.
"""
import math
import numpy as np
from .posenc_np import gen_sineembed_for_position

try:  # exact erf for GELU (nn.GELU() default = erf form, dab_transformer.py:84)
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf)


def linear(x, w, b):
    return x @ w.T + b


def relu(x):
    return np.maximum(x, 0)


def gelu(x):
    return (0.5 * x * (1.0 + _erf(x / np.sqrt(x.dtype.type(2.0))))).astype(x.dtype)


def layer_norm(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + x.dtype.type(eps)) * g + b


def softmax_last(x):
    # attention.py:400-401: softmax(x - max(x))
    x = x - x.max(-1, keepdims=True)
    e = np.exp(x)
    return e / e.sum(-1, keepdims=True)


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def inverse_sigmoid(x, eps=1e-5):
    # utils/misc.py:530-534
    x = np.clip(x, 0, 1)
    x1 = np.maximum(x, x.dtype.type(eps))
    x2 = np.maximum(1 - x, x.dtype.type(eps))
    return np.log(x1 / x2)


def mlp(x, W, prefix, n):
    # dab_transformer.py:45-48
    for i in range(n):
        x = linear(x, W[f"{prefix}.layers.{i}.weight"], W[f"{prefix}.layers.{i}.bias"])
        if i < n - 1:
            x = relu(x)
    return x


def mha_standard(q, k, v, nheads, w_o, b_o, key_padding_mask=None):
    """attention.py mode A (:336-341,377,409).  q [L,Nb,E], k [S,Nb,E], v [S,Nb,Ev] -> [L,Nb,Ev]."""
    L, Nb, E = q.shape
    S = k.shape[0]
    Ev = v.shape[2]
    hd, vd = E // nheads, Ev // nheads
    qh = (q * q.dtype.type(float(hd) ** -0.5)).reshape(L, Nb * nheads, hd).transpose(1, 0, 2)   # :293,336
    kh = k.reshape(S, Nb * nheads, hd).transpose(1, 0, 2)
    vh = v.reshape(S, Nb * nheads, vd).transpose(1, 0, 2)
    att = qh @ kh.transpose(0, 2, 1)                                                             # :377
    if key_padding_mask is not None:                                                             # :390-396
        att = att.reshape(Nb, nheads, L, S)
        att = np.where(key_padding_mask[:, None, None, :], -np.inf, att)
        att = att.reshape(Nb * nheads, L, S)
    att = softmax_last(att)
    out = att @ vh                                                                               # :409
    out = out.transpose(1, 0, 2).reshape(L, Nb, Ev)                                              # :414
    return linear(out, w_o, b_o)                                                                 # :415


def mha_query_specific(q, k, v, nheads, w_o, b_o, key_padding_mask=None):
    """attention.py mode B, query_specific_key (:343-346,379,411).
    q [nq,Nb,E], k [nq,S,Nb,E], v [nq,S,Nb,Ev] -> [nq,Nb,Ev]: every query row owns its keys/values."""
    nq, Nb, E = q.shape
    S = k.shape[1]
    Ev = v.shape[3]
    hd, vd = E // nheads, Ev // nheads
    qh = (q * q.dtype.type(float(hd) ** -0.5)).reshape(nq, Nb * nheads, hd).transpose(1, 0, 2)   # [Nb*H, nq, hd]
    kh = k.reshape(nq, S, Nb * nheads, hd).transpose(0, 2, 1, 3)                                 # [nq, Nb*H, S, hd]
    vh = v.reshape(nq, S, Nb * nheads, vd).transpose(0, 2, 1, 3)
    att = np.einsum("bnd,nbld->bnl", qh, kh)                                                     # :379
    if key_padding_mask is not None:
        att = att.reshape(Nb, nheads, nq, S)
        att = np.where(key_padding_mask[:, None, None, :], -np.inf, att)
        att = att.reshape(Nb * nheads, nq, S)
    att = softmax_last(att)
    out = np.einsum("bnl,nbld->bnd", att, vh)                                                    # :411
    out = out.transpose(1, 0, 2).reshape(nq, Nb, Ev)
    return linear(out, w_o, b_o)


def decoder_layer(W, p, tgt, memory, mask, pos, query_pos, query_sine_embed, is_first, nheads=8):
    """TransformerDecoderLayer.forward (dab_transformer.py:907-997).  Returns (tgt, tgt_temp, q_memory)."""
    g = lambda n: W[p + n]
    nq, BT, C = tgt.shape
    # self-attention :921-938
    q = linear(tgt, g("sa_qcontent_proj.weight"), g("sa_qcontent_proj.bias")) + \
        linear(query_pos, g("sa_qpos_proj.weight"), g("sa_qpos_proj.bias"))
    k = linear(tgt, g("sa_kcontent_proj.weight"), g("sa_kcontent_proj.bias")) + \
        linear(query_pos, g("sa_kpos_proj.weight"), g("sa_kpos_proj.bias"))
    v = linear(tgt, g("sa_v_proj.weight"), g("sa_v_proj.bias"))
    tgt2 = mha_standard(q, k, v, nheads, g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"))
    tgt = layer_norm(tgt + tgt2, g("norm1.weight"), g("norm1.bias"))
    # level-weighted, query-specific memory :943-946
    lvl_w = softmax_last(linear(tgt, g("lvl_w_embed.weight"), g("lvl_w_embed.bias")))            # [nq,BT,L]
    q_memory = np.einsum("ntl,lhtc->nhtc", lvl_w, memory)                                         # [nq,S,BT,C]
    q_memory = layer_norm(q_memory, g("norm_.weight"), g("norm_.bias"))
    # cross-attention :951-988
    q_content = linear(tgt, g("ca_qcontent_proj.weight"), g("ca_qcontent_proj.bias"))
    k_content = linear(q_memory, g("ca_kcontent_proj.weight"), g("ca_kcontent_proj.bias"))
    v = linear(q_memory, g("ca_v_proj.weight"), g("ca_v_proj.bias"))
    S = k_content.shape[1]
    k_pos = linear(pos[0:1], g("ca_kpos_proj.weight"), g("ca_kpos_proj.bias"))                    # :958 [1,S,BT,C]
    k_pos = np.broadcast_to(k_pos, (nq, S, BT, C))
    if is_first:                                                                                  # :964-970
        q = q_content + linear(query_pos, g("ca_qpos_proj.weight"), g("ca_qpos_proj.bias"))
        k = k_content + k_pos
    else:
        q, k = q_content, k_content
    hd = C // nheads
    q = q.reshape(nq, BT, nheads, hd)
    qse = linear(query_sine_embed, g("ca_qpos_sine_proj.weight"), g("ca_qpos_sine_proj.bias"))
    qse = qse.reshape(nq, BT, nheads, hd)
    q = np.concatenate([q, qse], axis=3).reshape(nq, BT, 2 * C)                                   # :975
    k = k.reshape(nq, S, BT, nheads, hd)
    k_pos = k_pos.reshape(nq, S, BT, nheads, hd)
    k = np.concatenate([k, k_pos], axis=4).reshape(nq, S, BT, 2 * C)                              # :979
    tgt2 = mha_query_specific(q, k, v, nheads, g("cross_attn.out_proj.weight"), g("cross_attn.out_proj.bias"),
                              key_padding_mask=mask)
    tgt = layer_norm(tgt + tgt2, g("norm2.weight"), g("norm2.bias"))
    tgt_temp = tgt
    tgt2 = linear(relu(linear(tgt, g("linear1.weight"), g("linear1.bias"))), g("linear2.weight"), g("linear2.bias"))
    tgt = layer_norm(tgt + tgt2, g("norm3.weight"), g("norm3.bias"))
    return tgt, tgt_temp, q_memory


def conv3x3(x, w, b):
    """x [N,C,H,W], w [O,C,3,3], padding 1 (dab_transformer.py:81,90) as im2col + GEMM."""
    N, C, H, Wd = x.shape
    xp = np.zeros((N, C, H + 2, Wd + 2), dtype=x.dtype)
    xp[:, :, 1:-1, 1:-1] = x
    cols = np.empty((N, H, Wd, C, 3, 3), dtype=x.dtype)
    for dy in range(3):
        for dx in range(3):
            cols[:, :, :, :, dy, dx] = xp[:, :, dy:dy + H, dx:dx + Wd].transpose(0, 2, 3, 1)
    y = cols.reshape(N * H * Wd, C * 9) @ w.reshape(w.shape[0], C * 9).T + b
    return y.reshape(N, H, Wd, w.shape[0])  # NHWC


def conv_block(W, p, x_nhwc):
    """ConvBlock.forward (dab_transformer.py:88-98), on NHWC data (the permutes at :91,96 are layout only)."""
    g = lambda n: W[p + n]
    x_nchw = x_nhwc.transpose(0, 3, 1, 2)
    y = conv3x3(x_nchw, g("conv1.weight"), g("conv1.bias"))
    y = layer_norm(y, g("norm.weight"), g("norm.bias"), eps=1e-6)                                 # :82
    y = linear(y, g("conv2.weight"), g("conv2.bias"))
    y = gelu(y)
    y = linear(y, g("conv3.weight"), g("conv3.bias"))
    return x_nhwc + y


def class_decoder_layer(W, p, actor_feature, q_memory, pos0, query_sine_embed, class_queries, orig_res, is_first,
                        nheads=8, taps=None):
    """TransformerClassDecoderLayer.forward (dab_transformer.py:1040-1079).
    actor_feature [nq,BT,C]; q_memory [nq,S,BT,C]; pos0 [S,BT,C]; query_sine_embed [nq,BT,C];
    class_queries [K,C] (first layer) or [K,N,C].  Returns (cls_output [nq,BT,K,C], next_query [K,N,C])."""
    g = lambda n: W[p + n]
    nq, BT, C = actor_feature.shape
    h, w = orig_res
    S = h * w
    N = nq * BT
    a2 = linear(relu(linear(actor_feature, g("cls_linear1.weight"), g("cls_linear1.bias"))),
                g("cls_linear2.weight"), g("cls_linear2.bias"))
    actor = layer_norm(actor_feature + a2, g("cls_norm.weight"), g("cls_norm.bias"))             # :1043-1045
    # :1049-1054  cls_feature[(n,b), y, x, :] = conv_norm(actor[n,b] + q_memory[n, y*w+x, b])
    enc = q_memory.transpose(0, 2, 1, 3).reshape(N, h, w, C)                                      # (N BT) H W D
    cls_feature = layer_norm(actor.reshape(N, 1, 1, C) + enc, g("conv_norm.weight"), g("conv_norm.bias"))
    if taps is not None:
        taps["cls_feature0"] = cls_feature
    for _ in range(3):                                                                            # :1055-1056
        cls_feature = conv_block(W, p + "conv_blocks.0.", cls_feature)
    if taps is not None:
        taps["cls_feature3"] = cls_feature
    # class-query self-attention :1059-1065
    if is_first:
        query = np.broadcast_to(class_queries[:, None, :], (class_queries.shape[0], N, C))
    else:
        query = class_queries
    K = query.shape[0]
    query2 = mha_standard(query, query, query, nheads, g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias"))
    query = layer_norm(query + query2, g("norm1.weight"), g("norm1.bias"))
    # cross-attention :1067-1071
    kx = linear(cls_feature.reshape(N, S, C), g("k_proj.weight").reshape(C, C), g("k_proj.bias"))      # 1x1 conv
    kx = kx.transpose(1, 0, 2)                                                                         # [S,N,C]
    pos_n = np.broadcast_to(pos0[:, None, :, :], (S, nq, BT, C)).reshape(S, N, C)
    key = np.concatenate([kx, pos_n], axis=-1)                                                         # [S,N,2C]
    cqp = linear(query_sine_embed, g("cls_qpos_sine_proj.weight"), g("cls_qpos_sine_proj.bias")).reshape(N, C)
    query_cat = np.concatenate([query, np.broadcast_to(cqp[None], (K, N, C))], axis=-1)               # [K,N,2C]
    value = linear(enc.reshape(N, S, C), g("v_proj.weight").reshape(C, C), g("v_proj.bias")).transpose(1, 0, 2)
    out = mha_standard(query_cat, key, value, nheads, g("cross_attn.out_proj.weight"), g("cross_attn.out_proj.bias"))
    cls_output = out.reshape(K, nq, BT, C).transpose(1, 2, 0, 3)                                      # [nq,BT,K,C]
    # FFN :1074-1077
    c2 = linear(relu(linear(cls_output, g("cls_linear1_.weight"), g("cls_linear1_.bias"))),
                g("cls_linear2_.weight"), g("cls_linear2_.bias"))
    cls_output = layer_norm(cls_output + c2, g("cls_norm_.weight"), g("cls_norm_.bias"))
    next_query = cls_output.transpose(2, 0, 1, 3).reshape(K, N, C)
    return cls_output, next_query


def decoder_forward(W, tgt, memory, mask, pos, refpoints_unsigmoid, orig_res, layers, dt=np.float32, taps=None):
    """TransformerDecoder.forward (dab_transformer.py:722-852) with bbox_embed attached (model.py:100-101),
    modulate_hw_attn=True, query_scale_type='cond_elewise', keep_query_pos=False, bbox_embed shared.
    Returns hs [Lr,BT,nq,C], cls_hs [Lr,BT,nq,K,C], references [Lr,BT,nq,4]."""
    W = {k: v.astype(dt) for k, v in W.items()}
    tgt, memory, pos, ref_u = (a.astype(dt) for a in (tgt, memory, pos, refpoints_unsigmoid))
    C = tgt.shape[-1]
    output = tgt
    reference_points = sigmoid(ref_u)                                                             # :735
    ref_points = [reference_points]
    class_queries = W["class_queries.weight"]
    inter, cls_inter = [], []
    for lid in range(layers):
        obj_center = reference_points[..., :4]
        qse_full = gen_sineembed_for_position(obj_center, dtype=dt)                               # :744
        query_pos = mlp(qse_full, W, "ref_point_head", 2)                                         # :745
        pos_tr = 1 if lid == 0 else mlp(output, W, "query_scale", 2)                              # :749-752
        qse = qse_full[..., :C] * pos_tr                                                          # :757
        refHW = sigmoid(mlp(output, W, "ref_anchor_head", 2))                                     # :761
        qse = qse.copy()
        qse[..., C // 2:] *= (refHW[..., 0] / obj_center[..., 2])[..., None]                      # :762
        qse[..., :C // 2] *= (refHW[..., 1] / obj_center[..., 3])[..., None]                      # :763
        output, actor, q_memory = decoder_layer(W, f"layers.{lid}.", output, memory, mask, pos, query_pos, qse,
                                                lid == 0)
        t = {} if taps is not None else None
        cls_output, class_queries = class_decoder_layer(W, f"cls_layers.{lid}.", actor, q_memory, pos[0], qse,
                                                        class_queries, orig_res, lid == 0, taps=t)
        if taps is not None:
            taps[f"l{lid}.output"] = output; taps[f"l{lid}.actor"] = actor; taps[f"l{lid}.q_memory"] = q_memory
            taps[f"l{lid}.qse"] = qse; taps[f"l{lid}.cls_output"] = cls_output
            for k, v in t.items():
                taps[f"l{lid}.{k}"] = v
        tmp = mlp(output, W, "bbox_embed", 3)                                                     # :817
        tmp = tmp[..., :4] + inverse_sigmoid(reference_points)                                    # :819
        new_ref = sigmoid(tmp)                                                                    # :820
        if lid != layers - 1:
            ref_points.append(new_ref)
        reference_points = new_ref                                                                # :823 (detach)
        inter.append(layer_norm(output, W["norm.weight"], W["norm.bias"]))                        # :826
        cls_inter.append(layer_norm(cls_output, W["cls_norm2.weight"], W["cls_norm2.bias"]))      # :827
    hs = np.stack(inter).transpose(0, 2, 1, 3)                                                    # :841
    cls_hs = np.stack(cls_inter).transpose(0, 2, 1, 3, 4)                                         # :842
    refs = np.stack(ref_points).transpose(0, 2, 1, 3)                                             # :843
    return hs, cls_hs, refs


def detr_heads(W, hs, cls_hs, reference):
    """models/model.py:191-221 (eval: dropout(0.5) is identity; shared bbox_embed).
    Returns pred_logits [Lr,BT,nq,K], pred_boxes [Lr,BT,nq,4], pred_logits_b [Lr,BT,nq,3]."""
    dt = hs.dtype
    W = {k: v.astype(dt) for k, v in W.items()}
    logits_b = linear(hs, W["heads.class_embed_b.weight"], W["heads.class_embed_b.bias"])         # :192
    tmp = mlp(hs, W, "bbox_embed", 3)                                                             # :198
    tmp = tmp[..., :4] + inverse_sigmoid(reference)                                               # :197,199
    boxes = sigmoid(tmp)                                                                          # :200
    logits = cls_hs.mean(-1)                                                                      # :219-221
    return logits, boxes, logits_b
