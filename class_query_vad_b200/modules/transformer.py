"""Drop-in for the reference `Transformer` (models/detr/dab_transformer.py:100-397) in its shipped configuration
(deformable 'attention' encoder, class-query decoder, no two-stage): same parameter names (`level_embed`, `encoder.layers.*`,
`decoder.*`: a reference state_dict loads with strict=True) and the same forward signature.  Every activation-sized step runs
in libcqvad.so: cqvad_level_to_tokens (flatten + level embedding), cqvad_deform_encoder_layer_forward, cqvad_encoder_to_decoder_memory
(un-flatten + make_interpolated_features + key-frame slice + rearrange) and the decoder engine.  Mask bookkeeping (valid ratios,
the flattened padding mask) is the reference's own few lines on boolean tensors."""
import torch
from torch import nn

from .. import _lib
from .decoder import build_decoder
from .encoder import DeformableTransformerEncoderLayer, DeformableTransformerEncoder, encoder_to_decoder_memory


def _flatten_levels_raw(srcs, pos_embeds, level_embed):
    lib = _lib.lib()
    dt, dev = srcs[0].dtype, srcs[0].device
    B = srcs[0].shape[0]
    shapes = [tuple(int(v) for v in s.shape[2:]) for s in srcs]
    ns = [t * h * w for t, h, w in shapes]
    Len = sum(ns)
    src_flat = torch.empty((B, Len, 256), dtype=dt, device=dev)
    pos_flat = torch.empty((B, Len, 256), dtype=dt, device=dev)
    le = level_embed.detach().to(device=dev, dtype=torch.float32).contiguous()
    p = _lib.ptr
    start = 0
    for l, (s, pe) in enumerate(zip(srcs, pos_embeds)):
        if s.shape[1] != 256:
            raise ValueError("d_model must be 256")
        sc, pc = s.detach().contiguous(), pe.detach().to(dt).contiguous()
        _lib.check(lib.cqvad_level_to_tokens(_lib.dtype_id(dt), p(sc), None, p(src_flat), B, ns[l], Len, start, _lib.stream_ptr()))
        _lib.check(lib.cqvad_level_to_tokens(_lib.dtype_id(dt), p(pc), p(le[l]), p(pos_flat), B, ns[l], Len, start, _lib.stream_ptr()))
        start += ns[l]
    return src_flat, pos_flat, shapes, ns


class LevelsToTokensFunction(torch.autograd.Function):
    """apply(level_embed, *srcs, *pos_embeds) -> (src_flatten, lvl_pos_embed_flatten); the backward (cqvad_level_to_tokens_backward)
    returns d level_embed [L, 256] (sum of the position-embedding gradient over clips and positions of each level) and the
    channel-first gradients of the levels (and of the position embeddings when they require grad)."""

    @staticmethod
    def forward(ctx, level_embed, *levels):
        L = len(levels) // 2
        srcs, pos_embeds = levels[:L], levels[L:]
        src_flat, pos_flat, shapes, ns = _flatten_levels_raw(srcs, pos_embeds, level_embed)
        ctx.meta = (L, [tuple(s.shape) for s in srcs], ns, [s.dtype for s in srcs], [pe.dtype for pe in pos_embeds], level_embed.dtype)
        ctx.need = [t.requires_grad for t in levels]
        return src_flat, pos_flat

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_src, g_pos):
        L, shp, ns, sdt, pdt, ledt = ctx.meta
        lib = _lib.lib()
        p = _lib.ptr
        dt, dev = g_src.dtype, g_src.device
        B, Len = g_src.shape[0], g_src.shape[1]
        g_src, g_pos = g_src.contiguous(), g_pos.contiguous()
        g_le = torch.zeros((L, 256), dtype=torch.float32, device=dev)
        outs_s, outs_p = [], []
        start = 0
        for l in range(L):
            gx = gp = None
            if ctx.need[l]:
                gx = torch.empty(shp[l], dtype=dt, device=dev)
                _lib.check(lib.cqvad_level_to_tokens_backward(_lib.dtype_id(dt), p(g_src), p(gx), None, B, ns[l], Len, start,
                                                              _lib.stream_ptr()))
            if ctx.need[L + l]:
                gp = torch.empty(shp[l], dtype=dt, device=dev)
            _lib.check(lib.cqvad_level_to_tokens_backward(_lib.dtype_id(dt), p(g_pos), p(gp), p(g_le[l]), B, ns[l], Len, start,
                                                          _lib.stream_ptr()))
            outs_s.append(None if gx is None else gx.to(sdt[l]))
            outs_p.append(None if gp is None else gp.to(pdt[l]))
            start += ns[l]
        return (g_le.to(ledt), *outs_s, *outs_p)


def flatten_levels(srcs, pos_embeds, level_embed):
    """dab_transformer.py:310-327: per level [B, 256, T, H, W] -> (src_flatten, lvl_pos_embed_flatten) [B, Len, 256], shapes
    [L, 3], level_start [L] through cqvad_level_to_tokens (differentiable: LevelsToTokensFunction)."""
    _lib.require_cuda(*srcs)
    dev = srcs[0].device
    shapes = [tuple(int(v) for v in s.shape[2:]) for s in srcs]
    if torch.is_grad_enabled() and (level_embed.requires_grad or any(t.requires_grad for t in (*srcs, *pos_embeds))):
        src_flat, pos_flat = LevelsToTokensFunction.apply(level_embed, *srcs, *pos_embeds)
    else:
        src_flat, pos_flat, _, _ = _flatten_levels_raw(srcs, pos_embeds, level_embed)
    sh, ls = _lib.shape_tensors(shapes, dev)
    return src_flat, pos_flat, sh, ls


class Transformer(nn.Module):
    def __init__(self, d_model=256, nhead=8, num_queries=15, num_encoder_layers=6, num_decoder_layers=6, dim_feedforward=2048,
                 dropout=0.1, num_feature_levels=4, enc_n_points=8, num_classes=80, temp_len=16):
        super().__init__()
        if d_model != 256 or nhead != 8:
            raise ValueError("libcqvad Transformer: d_model 256, 8 heads (every shipped configuration)")
        layer = DeformableTransformerEncoderLayer(d_model, dim_feedforward, dropout, "relu", num_feature_levels, nhead, enc_n_points)
        self.encoder = DeformableTransformerEncoder(layer, num_encoder_layers)
        self.decoder = build_decoder(num_queries=num_queries, num_classes=num_classes, num_layers=num_decoder_layers,
                                     dim_feedforward=dim_feedforward, d_model=d_model, nhead=nhead, dropout=dropout, temp_len=temp_len)
        self.level_embed = nn.Parameter(torch.randn(num_feature_levels, d_model))
        self.temp_len = temp_len
        self.num_feature_levels = num_feature_levels
        self.d_model = d_model
        self.eff = True          # models/model.py:52 (single-frame decoding)

    @staticmethod
    def get_valid_ratio(mask):
        # dab_transformer.py:228-237
        _, T, H, W = mask.shape
        valid_T = torch.sum(~mask[:, :, 0, 0], 1)
        valid_H = torch.sum(~mask[:, 0, :, 0], 1)
        valid_W = torch.sum(~mask[:, 0, 0, :], 1)
        return torch.stack([valid_W.float() / W, valid_H.float() / H, valid_T.float() / T], -1)

    def forward(self, srcs, masks, pos_embeds, refpoint_embed=None):
        assert refpoint_embed is not None
        bs = srcs[0].shape[0]
        src_flat, pos_flat, sh, ls = flatten_levels(srcs, pos_embeds, self.level_embed)
        mask_flat = torch.cat([m.flatten(1) for m in masks], 1)
        valid_ratios = torch.stack([self.get_valid_ratio(m) for m in masks], 1)
        # the padding mask is always handed over (an all-False mask is a no-op in the kernels): testing it with .any() would
        # synchronise the stream once per forward
        memory = self.encoder(src_flat, sh, ls, valid_ratios, pos=pos_flat, padding_mask=mask_flat)
        L = self.num_feature_levels
        Tt, H, W = (int(v) for v in srcs[L - 2].shape[2:])
        mem_l, pos0 = encoder_to_decoder_memory(memory, pos_flat, sh, ls, num_frames=self.temp_len, eff=self.eff)
        # mask of level -2 repeated in time (:250-252), key-frame slice, "(B T) (H W)" (:393)
        m = masks[L - 2].repeat(1, self.temp_len // masks[L - 2].size(1), 1, 1)
        if self.eff:
            t = m.size(1)
            m = m[:, t // 2:t // 2 + 1]
        mask = m.flatten(2).flatten(0, 1)
        refp = refpoint_embed[:, None].expand(-1, bs, -1, -1).flatten(1, 2)          # :371
        tgt = torch.zeros((refp.shape[0], mask.shape[0], self.d_model), device=src_flat.device, dtype=src_flat.dtype)
        return self.decoder(tgt, mem_l, memory_key_padding_mask=mask, pos=pos0[None].expand(L, -1, -1, -1),
                            refpoints_unsigmoid=refp, orig_res=(H, W))


def input_proj_levels(feats, convs, norms):
    """models/model.py:162-170 for the non-ViT (CSN) configurations: input_proj[l] = Conv3d + GroupNorm(32, 256) per level, written
    straight into the encoder's token sequence.  feats: the backbone levels [B, C_in, T, H, W]; convs / norms: the nn.Conv3d /
    nn.GroupNorm modules of input_proj in order -- the first len(feats) are the 1x1x1 projections (cqvad_input_proj_1x1_gn), any
    further one is the extra stride-(1,2,2) kernel-3 level applied to the LAST backbone feature (cqvad_input_proj_3x3s2_gn; the
    reference feeds a second extra level from the first one's output, which needs the channel-first tensor and is not covered).
    Returns (src_flatten [B, Len, 256], shapes [L, 3], level_start [L])."""
    _lib.require_cuda(*feats)
    lib = _lib.lib()
    dt, dev = feats[0].dtype, feats[0].device
    B = feats[0].shape[0]
    if len(convs) > len(feats) + 1:
        raise ValueError("input_proj_levels: at most one extra (stride-2) level")
    shapes = [tuple(int(v) for v in f.shape[2:]) for f in feats]
    if len(convs) == len(feats) + 1:
        t, h, w = shapes[-1]
        shapes.append((t, (h - 1) // 2 + 1, (w - 1) // 2 + 1))
    ns = [t * h * w for t, h, w in shapes]
    Len = sum(ns)
    tokens = torch.empty((B, Len, 256), dtype=dt, device=dev)
    p = _lib.ptr
    f32 = lambda t_: None if t_ is None else t_.detach().to(device=dev, dtype=torch.float32).contiguous()
    start = 0
    for l, (conv, gn) in enumerate(zip(convs, norms)):
        if gn.num_groups != 32 or conv.out_channels != 256:
            raise ValueError("input_proj_levels covers Conv3d -> GroupNorm(32, 256)")
        g, be, b = f32(gn.weight), f32(gn.bias), f32(conv.bias)
        if l < len(feats):
            f = feats[l].contiguous()
            Cin = f.shape[1]
            if tuple(conv.kernel_size) != (1, 1, 1):
                raise ValueError("backbone levels are projected with kernel_size 1")
            w = conv.weight.detach().reshape(256, Cin).to(device=dev, dtype=dt).contiguous()
            need = lib.cqvad_input_proj_workspace_bytes(_lib.dtype_id(dt), B, Cin, ns[l])
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(lib.cqvad_input_proj_1x1_gn(_lib.dtype_id(dt), p(f), p(w), p(b), p(g), p(be), float(gn.eps), p(tokens), p(ws), need,
                                                   B, Cin, ns[l], Len, start, _lib.stream_ptr()))
        else:
            f = feats[-1].contiguous()
            Cin, (T, H, W) = f.shape[1], f.shape[2:]
            if tuple(conv.kernel_size) != (3, 3, 3) or tuple(conv.stride) != (1, 2, 2) or tuple(conv.padding) != (1, 1, 1):
                raise ValueError("the extra level is Conv3d(kernel_size=3, stride=(1, 2, 2), padding=1)")
            w = conv.weight.detach().permute(0, 2, 3, 4, 1).reshape(256, 27 * Cin).to(device=dev, dtype=dt).contiguous()
            need = lib.cqvad_input_proj_3x3s2_workspace_bytes(_lib.dtype_id(dt), B, Cin, T, H, W)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            _lib.check(lib.cqvad_input_proj_3x3s2_gn(_lib.dtype_id(dt), p(f), p(w), p(b), p(g), p(be), float(gn.eps), p(tokens), p(ws), need,
                                                     B, Cin, T, H, W, Len, start, _lib.stream_ptr()))
        start += ns[l]
    sh, ls = _lib.shape_tensors(shapes, dev)
    return tokens, sh, ls
