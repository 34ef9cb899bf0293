"""Drop-ins for the reference `DeformableTransformerEncoderLayer` / `DeformableTransformerEncoder`
(models/detr/dab_transformer.py:425-523): same constructors, forward signatures and parameter names (a reference state_dict
loads with strict=True); one layer = ONE C-ABI call, cqvad_deform_encoder_layer_forward (csrc/encoder.cu); with gradients
enabled the layer goes through EncoderLayerFunction (cqvad_deform_encoder_layer_train_forward / _backward).  In train() mode
dropout1/2/3 are applied there (Philox masks); no torch arithmetic on the activation path -- `get_reference_points` (a [B, Len, L, 3] meshgrid of the
level shapes and valid ratios, computed once per forward) is host-side glue written with torch like the reference's."""
import ctypes
import copy

import torch
from torch import nn

from .. import _lib
from .ms_deform_attn import MSDeformAttn3D

ENC_WEIGHT_ORDER = ("self_attn.sampling_offsets", "self_attn.attention_weights", "self_attn.value_proj", "self_attn.output_proj",
                    "norm1", "linear1", "linear2", "norm2")
_MATRICES = {"self_attn.sampling_offsets", "self_attn.attention_weights", "self_attn.value_proj", "self_attn.output_proj",
             "linear1", "linear2"}


def pack_encoder_layer_weights(state, dtype, device, prefix=""):
    """[(tensor kept alive)], ctypes pointer table in the order include/cqvad.h documents."""
    keep = []
    for base in ENC_WEIGHT_ORDER:
        for leaf in ("weight", "bias"):
            t = state[f"{prefix}{base}.{leaf}"]
            t = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)
            dt = dtype if (leaf == "weight" and base in _MATRICES) else torch.float32
            keep.append(t.detach().to(device=device, dtype=dt).contiguous())
    tab = (ctypes.c_void_p * len(keep))(*[k.data_ptr() for k in keep])
    return keep, tab


def encoder_layer_forward(packed, src, pos, reference_points, shapes, level_start, padding_mask, n_points, d_ffn, want_attn=False):
    """packed = pack_encoder_layer_weights(...).  src/pos [B, Len, 256]; returns out (and the attention module's output)."""
    keep, tab = packed
    _lib.require_cuda(src, pos, reference_points, shapes, level_start)
    lib = _lib.lib()
    dt = src.dtype
    B, Len, C = src.shape
    if C != 256:
        raise ValueError("d_model must be 256")
    L = int(shapes.shape[0])
    src_c, pos_c = src.contiguous(), pos.to(dt).contiguous()
    refp = reference_points.to(torch.float32).contiguous()
    if tuple(refp.shape) != (B, Len, L, 3):
        raise ValueError("reference_points must be [B, Len, n_levels, 3]")
    sh, ls = shapes.to(torch.int64).contiguous(), level_start.to(torch.int64).contiguous()
    m8 = None
    if padding_mask is not None:
        m8 = padding_mask.to(src.device).contiguous()
        m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else m8.to(torch.uint8)
    need = lib.cqvad_deform_encoder_layer_workspace_bytes(_lib.dtype_id(dt), B, Len, L, n_points, d_ffn)
    ws = torch.empty(need, dtype=torch.uint8, device=src.device)
    out = torch.empty_like(src_c)
    attn_out = torch.empty_like(src_c) if want_attn else None
    p = _lib.ptr
    _lib.check(lib.cqvad_deform_encoder_layer_forward(_lib.dtype_id(dt), tab, p(src_c), p(pos_c), p(refp), p(sh), p(ls), p(m8), p(out),
                                                      p(attn_out), p(ws), need, B, Len, L, n_points, d_ffn, _lib.stream_ptr()))
    return (out, attn_out) if want_attn else out


# Training workspaces (several GB per layer at the 33 320-token pyramid) are recycled through a small free list instead of going
# back to the caching allocator after every backward: a forward takes one, the matching backward returns it.
_TRAIN_WS_FREE = {}


def _take_train_ws(need, device):
    free = _TRAIN_WS_FREE.setdefault(str(device), [])
    for i, t in enumerate(free):
        if t.numel() >= need:
            return free.pop(i)
    return torch.empty(need, dtype=torch.uint8, device=device)


def _give_train_ws(ws):
    free = _TRAIN_WS_FREE.setdefault(str(ws.device), [])
    if len(free) < 16:
        free.append(ws)


class EncoderLayerFunction(torch.autograd.Function):
    """loss.backward() through one encoder layer: cqvad_deform_encoder_layer_train_forward / _backward.
    apply(src, pos, reference_points, shapes, level_start, padding_mask, n_points, d_ffn, *16 parameters in state_dict order)."""

    @staticmethod
    def forward(ctx, src, pos, reference_points, shapes, level_start, padding_mask, n_points, d_ffn, p_drop, seed, *params):
        _lib.require_cuda(src, pos, reference_points, shapes, level_start)
        lib = _lib.lib()
        dt = src.dtype
        B, Len, C = src.shape
        L = int(shapes.shape[0])
        names = [f"{b}.{l}" for b in ENC_WEIGHT_ORDER for l in ("weight", "bias")]
        keep, tab = pack_encoder_layer_weights(dict(zip(names, params)), dt, src.device)
        src_c, pos_c = src.detach().contiguous(), pos.detach().to(dt).contiguous()
        refp = reference_points.detach().to(torch.float32).contiguous()
        sh, ls = shapes.to(torch.int64).contiguous(), level_start.to(torch.int64).contiguous()
        m8 = None
        if padding_mask is not None:
            m8 = padding_mask.to(src.device).contiguous()
            m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else m8.to(torch.uint8)
        need = lib.cqvad_deform_encoder_layer_train_workspace_bytes(_lib.dtype_id(dt), B, Len, L, n_points, d_ffn)
        ws = _take_train_ws(need, src.device)
        out = torch.empty_like(src_c)
        p = _lib.ptr
        _lib.check(lib.cqvad_deform_encoder_layer_train_forward(_lib.dtype_id(dt), tab, p(src_c), p(pos_c), p(refp), p(sh), p(ls), p(m8),
                                                                p(out), p(ws), need, B, Len, L, n_points, d_ffn, float(p_drop), int(seed),
                                                                _lib.stream_ptr()))
        ctx.saved = (keep, tab, src_c, sh, ls, m8, ws, need, (B, Len, L, n_points, d_ffn, float(p_drop), int(seed)), [q.dtype for q in params])
        return out

    @staticmethod
    def backward(ctx, grad_out):
        keep, tab, src_c, sh, ls, m8, ws, need, (B, Len, L, n_points, d_ffn, p_drop, seed), pdt = ctx.saved
        lib = _lib.lib()
        dt = src_c.dtype
        go = grad_out.to(dt).contiguous()
        gsrc, gpos = torch.empty_like(src_c), torch.empty_like(src_c)
        gw = [torch.zeros(k.shape, dtype=torch.float32, device=src_c.device) for k in keep]
        gtab = (ctypes.c_void_p * len(gw))(*[g.data_ptr() for g in gw])
        p = _lib.ptr
        _lib.check(lib.cqvad_deform_encoder_layer_backward(_lib.dtype_id(dt), tab, p(src_c), p(sh), p(ls), p(m8), p(go), p(gsrc), p(gpos),
                                                           gtab, p(ws), need, B, Len, L, n_points, d_ffn, p_drop, seed, _lib.stream_ptr()))
        _give_train_ws(ws)
        ctx.saved = None
        return (gsrc, gpos, None, None, None, None, None, None, None, None) + tuple(g.to(d) for g, d in zip(gw, pdt))


class DeformableTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model=256, d_ffn=1024, dropout=0.1, activation="relu", n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model != 256 or n_heads != 8 or activation != "relu":
            raise ValueError("libcqvad encoder layer: d_model 256, 8 heads, relu (every shipped configuration)")
        self.self_attn = MSDeformAttn3D(d_model, n_levels, n_heads, n_points)
        self.dropout1 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.linear1 = nn.Linear(d_model, d_ffn)
        self.dropout2 = nn.Dropout(dropout)
        self.linear2 = nn.Linear(d_ffn, d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm2 = nn.LayerNorm(d_model)
        self.n_points, self.d_ffn = n_points, d_ffn
        self._packed = None
        self.dropout_seed = 0x4321            # Philox key of this layer's dropout masks (+ a per-call counter)

    def _pack(self, dtype, device):
        ver = tuple(p._version for p in self.parameters()) + (dtype, str(device))
        if self._packed is None or self._packed[0] != ver:
            self._packed = (ver, pack_encoder_layer_weights(self.state_dict(), dtype, device))
        return self._packed[1]

    def weights_updated(self):
        self._packed = None

    def forward(self, src, pos, reference_points, spatio_temporal_shapes, level_start_index, padding_mask=None):
        if pos is None:
            pos = torch.zeros_like(src)
        if torch.is_grad_enabled() and (src.requires_grad or any(p.requires_grad for p in self.parameters())):
            sd = dict(self.named_parameters())
            params = [sd[f"{b}.{l}"] for b in ENC_WEIGHT_ORDER for l in ("weight", "bias")]
            p_drop = float(self.dropout1.p) if self.training else 0.0          # dropout1 / 2 / 3 (dab_transformer.py:499-519)
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            seed = (int(getattr(self, "dropout_seed", 0x4321)) << 24) + self._train_calls
            return EncoderLayerFunction.apply(src, pos, reference_points, spatio_temporal_shapes, level_start_index, padding_mask,
                                              self.n_points, self.d_ffn, p_drop, seed, *params)
        return encoder_layer_forward(self._pack(src.dtype, src.device), src, pos, reference_points, spatio_temporal_shapes,
                                     level_start_index, padding_mask, self.n_points, self.d_ffn)


class DeformableTransformerEncoder(nn.Module):
    def __init__(self, encoder_layer, num_layers, gradient_checkpointing=False):
        super().__init__()
        self.layers = nn.ModuleList([copy.deepcopy(encoder_layer) for _ in range(num_layers)])
        for i, layer in enumerate(self.layers):
            layer.dropout_seed = 0x4321 + i          # distinct Philox keys per layer (a fresh stream per call on top)
        self.num_layers = num_layers
        self.gradient_checkpointing = gradient_checkpointing

    @staticmethod
    def get_reference_points(spatio_temporal_shapes, valid_ratios, device):
        """Reference points [B, Len, L, 3] of dab_transformer.py:433-449 (bit for bit: tests/test_reference_points_cpu.py): voxel
        centres (i + 0.5) / (valid_ratio * extent) per axis in (x, y, t) order, every level's points then rescaled by the valid
        ratios of ALL levels.  Built per axis and broadcast over the level's grid: the per-voxel divisions of the meshgrid
        formulation are the same divisions on replicated operands."""
        B = valid_ratios.shape[0]

        def centres(extent, ratio):                      # [B, extent]
            grid = torch.linspace(0.5, extent - 0.5, extent, dtype=torch.float32, device=device)
            return grid[None] / (ratio[:, None] * extent)

        per_level = []
        for lvl, (T_, H_, W_) in enumerate(_lib.host_shapes(spatio_temporal_shapes)):
            t = centres(T_, valid_ratios[:, lvl, 2])[:, :, None, None].expand(B, T_, H_, W_)
            y = centres(H_, valid_ratios[:, lvl, 1])[:, None, :, None].expand(B, T_, H_, W_)
            x = centres(W_, valid_ratios[:, lvl, 0])[:, None, None, :].expand(B, T_, H_, W_)
            per_level.append(torch.stack((x, y, t), -1).reshape(B, T_ * H_ * W_, 3))
        return torch.cat(per_level, 1)[:, :, None] * valid_ratios[:, None]

    def forward(self, src, spatio_temporal_shapes, level_start_index, valid_ratios, pos=None, padding_mask=None):
        output = src
        reference_points = self.get_reference_points(spatio_temporal_shapes, valid_ratios, device=src.device)
        for layer in self.layers:
            output = layer(output, pos, reference_points, spatio_temporal_shapes, level_start_index, padding_mask)
        return output


def _encoder_to_decoder_memory_raw(tokens, pos_tokens, shapes, level_start, num_frames, eff):
    _lib.require_cuda(tokens, shapes, level_start)
    dt = tokens.dtype
    B, Len, C = tokens.shape
    L = int(shapes.shape[0])
    Tt, H, W = (int(v) for v in _lib.host_shapes(shapes)[L - 2])
    Tp = 1 if eff else int(num_frames)
    tok = tokens.detach().contiguous()
    ptok = None if pos_tokens is None else pos_tokens.detach().to(dt).contiguous()
    sh, ls = shapes.to(torch.int64).contiguous(), level_start.to(torch.int64).contiguous()
    memory = torch.empty((L, H * W, B * Tp, C), dtype=dt, device=tokens.device)
    pos0 = None if ptok is None else torch.empty((H * W, B * Tp, C), dtype=dt, device=tokens.device)
    p = _lib.ptr
    _lib.check(_lib.lib().cqvad_encoder_to_decoder_memory(_lib.dtype_id(dt), p(tok), p(ptok), p(sh), p(ls), L, B, Len, Tt, H, W,
                                                         int(num_frames), 1 if eff else 0, p(memory), p(pos0), _lib.stream_ptr()))
    return memory, pos0, (sh, ls, L, B, Len, Tt, H, W)


class EncoderToDecoderMemoryFunction(torch.autograd.Function):
    """apply(tokens, pos_tokens, shapes, level_start, num_frames, eff) -> (memory, pos0).  backward: the trilinear scatter of
    d(memory) into the token gradient (cqvad_encoder_to_decoder_memory_backward).  pos0 is marked non-differentiable: the decoder
    consumes it only on the key side of softmax attentions."""

    @staticmethod
    def forward(ctx, tokens, pos_tokens, shapes, level_start, num_frames, eff):
        memory, pos0, geo = _encoder_to_decoder_memory_raw(tokens, pos_tokens, shapes, level_start, num_frames, eff)
        ctx.geo, ctx.nf, ctx.eff, ctx.dt = geo, int(num_frames), bool(eff), tokens.dtype
        if pos0 is not None:
            ctx.mark_non_differentiable(pos0)
        return memory, pos0

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_mem, _g_pos0):
        sh, ls, L, B, Len, Tt, H, W = ctx.geo
        dt = ctx.dt
        gm = g_mem.to(torch.float32).contiguous()
        gtok = torch.empty((B, Len, 256), dtype=dt, device=gm.device)
        ws = None if dt == torch.float32 else torch.empty((B, Len, 256), dtype=torch.float32, device=gm.device)
        p = _lib.ptr
        _lib.check(_lib.lib().cqvad_encoder_to_decoder_memory_backward(_lib.dtype_id(dt), p(gm), p(sh), p(ls), L, B, Len, Tt, H, W, ctx.nf,
                                                                      1 if ctx.eff else 0, p(gtok), p(ws), _lib.stream_ptr()))
        return gtok, None, None, None, None, None


def encoder_to_decoder_memory(tokens, pos_tokens, shapes, level_start, num_frames, eff=True):
    """The encoder -> decoder step of Transformer.forward (dab_transformer.py:349-393) through cqvad_encoder_to_decoder_memory:
    tokens / pos_tokens [B, Len, 256] -> (memory [L, H*W, B*T', 256], pos0 [H*W, B*T', 256]) with (T, H, W) = shapes[-2].
    Differentiable with respect to `tokens` (EncoderToDecoderMemoryFunction)."""
    if torch.is_grad_enabled() and tokens.requires_grad:
        return EncoderToDecoderMemoryFunction.apply(tokens, pos_tokens, shapes, level_start, num_frames, eff)
    memory, pos0, _ = _encoder_to_decoder_memory_raw(tokens, pos_tokens, shapes, level_start, num_frames, eff)
    return memory, pos0
