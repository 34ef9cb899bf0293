"""TEST INFRASTRUCTURE -- torch (CPU) restatement of the reference class-query decoder, used where GRADIENTS are needed on
the host: the cpu_baseline / `--impl reference` legs of bench.py for the training step (BASELINE.json configs[1]: decoder
fwd + bwd) and the CPU tests that pin it.  Same structure and reference citations as oracle/decoder_np.py (the numpy
restatement), written with the torch functional ops the reference itself calls (F.linear / layer_norm / conv2d / gelu /
softmax: SURVEY.md section 8c "Third-party arithmetic"), so that torch autograd yields the backward the reference training
loop runs (train.py:151 loss.backward()).  Never imported by the product package.

Pinned by tests/test_oracle_golden.py against the reference-generated fixtures: forward (tests/golden/dec_*.npz) and
gradients (tests/golden/grad_*.npz, produced by autograd on the UNMODIFIED reference, oracle/make_golden_grads.py).
"""
import torch
import torch.nn.functional as F


def _mlp(x, W, prefix, n):                                   # dab_transformer.py:45-48
    for i in range(n):
        x = F.linear(x, W[f"{prefix}.layers.{i}.weight"], W[f"{prefix}.layers.{i}.bias"])
        if i < n - 1:
            x = F.relu(x)
    return x


def _inverse_sigmoid(x, eps=1e-5):                           # utils/misc.py:530-534
    x = x.clamp(min=0, max=1)
    return torch.log(x.clamp(min=eps) / (1 - x).clamp(min=eps))


def gen_sineembed_for_position(pos):                         # dab_transformer.py:50-76
    import math
    scale = 2 * math.pi
    dim_t = torch.arange(128, dtype=pos.dtype)
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / 128)
    out = []
    for idx in (1, 0, 2, 3):                                 # (y, x, w, h)
        e = pos[..., idx, None] * scale / dim_t
        out.append(torch.stack((e[..., 0::2].sin(), e[..., 1::2].cos()), dim=-1).flatten(-2))
    return torch.cat(out, dim=-1)


def _softmax(x):                                             # attention.py:400-401
    return torch.softmax(x - x.max(-1, keepdim=True)[0], dim=-1)


def mha_standard(q, k, v, H, w_o, b_o, kpm=None):            # attention.py mode A (:336-341,377,409)
    L, Nb, E = q.shape
    S, Ev = k.shape[0], v.shape[2]
    hd, vd = E // H, Ev // H
    qh = (q * float(hd) ** -0.5).reshape(L, Nb * H, hd).transpose(0, 1)
    kh = k.reshape(S, Nb * H, hd).transpose(0, 1)
    vh = v.reshape(S, Nb * H, vd).transpose(0, 1)
    att = torch.bmm(qh, kh.transpose(1, 2))
    if kpm is not None:
        att = att.view(Nb, H, L, S).masked_fill(kpm[:, None, None, :], float("-inf")).view(Nb * H, L, S)
    out = torch.bmm(_softmax(att), vh).transpose(0, 1).reshape(L, Nb, Ev)
    return F.linear(out, w_o, b_o)


def mha_query_specific(q, k, v, H, w_o, b_o, kpm=None):      # attention.py mode B (:343-346,379,411)
    nq, Nb, E = q.shape
    S, Ev = k.shape[1], v.shape[3]
    hd, vd = E // H, Ev // H
    qh = (q * float(hd) ** -0.5).reshape(nq, Nb * H, hd).transpose(0, 1)
    kh = k.reshape(nq, S, Nb * H, hd).permute(0, 2, 1, 3)
    vh = v.reshape(nq, S, Nb * H, vd).permute(0, 2, 1, 3)
    att = torch.einsum("bnd,nbld->bnl", qh, kh)
    if kpm is not None:
        att = att.view(Nb, H, nq, S).masked_fill(kpm[:, None, None, :], float("-inf")).view(Nb * H, nq, S)
    out = torch.einsum("bnl,nbld->bnd", _softmax(att), vh).transpose(0, 1).reshape(nq, Nb, Ev)
    return F.linear(out, w_o, b_o)


def decoder_layer(W, p, tgt, memory, mask, pos, query_pos, qse, is_first, H=8, drop=None):     # dab_transformer.py:907-997
    drop = drop or (lambda x, k: x)          # dropout hook: (tensor, site 0..3) -> tensor  (:937,991,995)
    g = lambda n: W[p + n]
    lin = lambda x, n: F.linear(x, g(n + ".weight"), g(n + ".bias"))
    nq, BT, C = tgt.shape
    q = lin(tgt, "sa_qcontent_proj") + lin(query_pos, "sa_qpos_proj")
    k = lin(tgt, "sa_kcontent_proj") + lin(query_pos, "sa_kpos_proj")
    v = lin(tgt, "sa_v_proj")
    tgt = F.layer_norm(tgt + drop(mha_standard(q, k, v, H, g("self_attn.out_proj.weight"), g("self_attn.out_proj.bias")), 0), (C,),
                       g("norm1.weight"), g("norm1.bias"))
    lvl_w = lin(tgt, "lvl_w_embed").softmax(-1)
    q_memory = F.layer_norm(torch.einsum("ntl,lhtc->nhtc", lvl_w, memory), (C,), g("norm_.weight"), g("norm_.bias"))
    q_content = lin(tgt, "ca_qcontent_proj")
    k_content = lin(q_memory, "ca_kcontent_proj")
    v = lin(q_memory, "ca_v_proj")
    S = k_content.shape[1]
    k_pos = lin(pos, "ca_kpos_proj")[0:1].expand(nq, -1, -1, -1)
    if is_first:
        q = q_content + lin(query_pos, "ca_qpos_proj")
        k = k_content + k_pos
    else:
        q, k = q_content, k_content
    hd = C // H
    q = torch.cat([q.view(nq, BT, H, hd), lin(qse, "ca_qpos_sine_proj").view(nq, BT, H, hd)], dim=3).view(nq, BT, 2 * C)
    k = torch.cat([k.view(nq, S, BT, H, hd), k_pos.reshape(nq, S, BT, H, hd)], dim=4).view(nq, S, BT, 2 * C)
    tgt2 = mha_query_specific(q, k, v, H, g("cross_attn.out_proj.weight"), g("cross_attn.out_proj.bias"), kpm=mask)
    tgt = F.layer_norm(tgt + drop(tgt2, 1), (C,), g("norm2.weight"), g("norm2.bias"))
    tgt_temp = tgt
    tgt = F.layer_norm(tgt + drop(lin(drop(F.relu(lin(tgt, "linear1")), 2), "linear2"), 3), (C,), g("norm3.weight"), g("norm3.bias"))
    return tgt, tgt_temp, q_memory


def conv_block(W, p, x):                                     # dab_transformer.py:88-98, x NCHW
    g = lambda n: W[p + n]
    y = F.conv2d(x, g("conv1.weight"), g("conv1.bias"), padding=1).permute(0, 2, 3, 1)
    y = F.layer_norm(y, (y.shape[-1],), g("norm.weight"), g("norm.bias"), eps=1e-6)
    y = F.linear(F.gelu(F.linear(y, g("conv2.weight"), g("conv2.bias"))), g("conv3.weight"), g("conv3.bias"))
    return x + y.permute(0, 3, 1, 2)


def class_decoder_layer(W, p, actor_feature, q_memory, pos0, qse, class_queries, orig_res, is_first, H=8, drop=None):   # :1040-1079
    drop = drop or (lambda x, k: x)          # dropout hook, sites 4..8 (:1043-1044,1062,1076-1077)
    g = lambda n: W[p + n]
    lin = lambda x, n: F.linear(x, g(n + ".weight"), g(n + ".bias"))
    nq, BT, C = actor_feature.shape
    h, w = orig_res
    S, N = h * w, nq * BT
    actor = F.layer_norm(actor_feature + drop(lin(drop(F.relu(lin(actor_feature, "cls_linear1")), 4), "cls_linear2"), 5), (C,),
                         g("cls_norm.weight"), g("cls_norm.bias"))
    enc = q_memory.permute(0, 2, 3, 1).reshape(N, C, h, w)                                        # (N BT) D H W
    feat = actor.reshape(N, C, 1, 1) + enc
    feat = F.layer_norm(feat.permute(0, 2, 3, 1), (C,), g("conv_norm.weight"), g("conv_norm.bias")).permute(0, 3, 1, 2)
    for _ in range(3):
        feat = conv_block(W, p + "conv_blocks.0.", feat)
    query = class_queries[:, None].expand(-1, N, -1) if is_first else class_queries
    K = query.shape[0]
    query = F.layer_norm(query + drop(mha_standard(query, query, query, H, g("self_attn.out_proj.weight"),
                                                   g("self_attn.out_proj.bias")), 6), (C,), g("norm1.weight"), g("norm1.bias"))
    kx = F.conv2d(feat, g("k_proj.weight"), g("k_proj.bias")).flatten(2).permute(2, 0, 1)
    key = torch.cat([kx, pos0[:, None].expand(-1, nq, -1, -1).flatten(1, 2)], dim=-1)
    cqp = lin(qse, "cls_qpos_sine_proj").flatten(0, 1)[None].expand(K, -1, -1)
    value = F.conv2d(enc, g("v_proj.weight"), g("v_proj.bias")).flatten(2).permute(2, 0, 1)
    out = mha_standard(torch.cat([query, cqp], dim=-1), key, value, H, g("cross_attn.out_proj.weight"),
                       g("cross_attn.out_proj.bias"))
    cls_output = out.reshape(K, nq, BT, C).permute(1, 2, 0, 3)
    cls_output = F.layer_norm(cls_output + drop(lin(drop(F.relu(lin(cls_output, "cls_linear1_")), 7), "cls_linear2_"), 8), (C,),
                              g("cls_norm_.weight"), g("cls_norm_.bias"))
    return cls_output, cls_output.permute(2, 0, 1, 3).flatten(1, 2)


def decoder_forward(W, tgt, memory, mask, pos, refpoints_unsigmoid, orig_res, layers, drop=None):
    """TransformerDecoder.forward (dab_transformer.py:722-852) on torch CPU tensors (W: name -> tensor, possibly with
    requires_grad).  Returns hs [Lr,BT,nq,C], cls_hs [Lr,BT,nq,K,C], references [Lr,BT,nq,4]."""
    C = tgt.shape[-1]
    output = tgt
    reference_points = refpoints_unsigmoid.sigmoid()
    ref_points = [reference_points]
    class_queries = W["class_queries.weight"]
    inter, cls_inter = [], []
    for lid in range(layers):
        obj_center = reference_points[..., :4]
        qse_full = gen_sineembed_for_position(obj_center)
        query_pos = _mlp(qse_full, W, "ref_point_head", 2)
        pos_tr = 1 if lid == 0 else _mlp(output, W, "query_scale", 2)
        qse = qse_full[..., :C] * pos_tr
        refHW = _mlp(output, W, "ref_anchor_head", 2).sigmoid()
        qse = torch.cat([qse[..., :C // 2] * (refHW[..., 1] / obj_center[..., 3]).unsqueeze(-1),
                         qse[..., C // 2:] * (refHW[..., 0] / obj_center[..., 2]).unsqueeze(-1)], dim=-1)   # :762-763
        dl = None if drop is None else (lambda x, k, _l=lid: drop(x, _l, k))          # training-mode nn.Dropout sites of this layer pair
        output, actor, q_memory = decoder_layer(W, f"layers.{lid}.", output, memory, mask, pos, query_pos, qse, lid == 0, drop=dl)
        cls_output, class_queries = class_decoder_layer(W, f"cls_layers.{lid}.", actor.clone().detach(), q_memory, pos[0], qse,
                                                        class_queries, orig_res, lid == 0, drop=dl)          # :810
        tmp = _mlp(output, W, "bbox_embed", 3)
        new_ref = (tmp[..., :4] + _inverse_sigmoid(reference_points)).sigmoid()
        if lid != layers - 1:
            ref_points.append(new_ref)
        reference_points = new_ref.detach()                                                          # :823
        inter.append(F.layer_norm(output, (C,), W["norm.weight"], W["norm.bias"]))
        cls_inter.append(F.layer_norm(cls_output, (C,), W["cls_norm2.weight"], W["cls_norm2.bias"]))
    return (torch.stack(inter).transpose(1, 2), torch.stack(cls_inter).transpose(1, 2), torch.stack(ref_points).transpose(1, 2))


def train_step(Wnp, inp, lw, layers, threads=None, drop=None):
    """One decoder fwd + bwd on the host: loss = sum(w_hs*hs) + sum(w_cls*cls_hs) + sum(w_refs*refs).  Returns
    (loss, {name: grad ndarray}, grad_memory, grad_tgt, grad_refpoints)."""
    if threads:
        torch.set_num_threads(threads)
    t = lambda a: torch.from_numpy(a.copy())
    W = {k: t(v).requires_grad_(True) for k, v in Wnp.items() if not k.startswith("heads.") and ".conv_blocks.1." not in k
         and ".conv_blocks.2." not in k}
    tgt, memory, ref = (t(inp[k]).requires_grad_(True) for k in ("tgt", "memory", "refpoints_unsigmoid"))
    hs, cls_hs, refs = decoder_forward(W, tgt, memory, t(inp["mask"]), t(inp["pos"]), ref, inp["orig_res"], layers, drop=drop)
    loss = (t(lw["w_hs"]) * hs).sum() + (t(lw["w_cls"]) * cls_hs).sum() + (t(lw["w_refs"]) * refs).sum()
    loss.backward()
    grads = {k: (torch.zeros_like(v) if v.grad is None else v.grad).numpy() for k, v in W.items()}
    return float(loss.item()), grads, memory.grad.numpy(), tgt.grad.numpy(), ref.grad.numpy()
