"""Drop-ins for the decoder classes of models/detr/dab_transformer.py: MLP (:36-48), ConvBlock (:78-98),
TransformerDecoderLayer (:854-997), TransformerClassDecoderLayer (:999-1079), TransformerDecoder (:671-852).

Constructors, attribute/parameter names and forward signatures are the reference's, so `load_state_dict` of a
reference checkpoint works and `models/model.py:100-101,191` (which injects `decoder.bbox_embed` and calls the
transformer) is unchanged.  `TransformerDecoder.forward` runs the whole layer stack natively (one call of
cqvad_decoder_forward; with gradients enabled cqvad_decoder_train_forward / _backward behind DecoderFunction, where the
nn.Dropout modules' p is applied in train() mode).  The per-layer `forward`s are stand-alone inference conveniences: they
compose the building-block entry points (linear, LayerNorm, attention core, ConvBlock) and keep small tensor glue in torch
(the 4-way level softmax + einsum mix, per-head concatenations); they are not on the measured path and apply no dropout.
"""
import copy

import torch
from torch import nn

from .. import _lib
from ..engine import DecoderEngine
from .attention import MultiheadAttention
from .ops import linear, layer_norm, conv_block
from .position_encoding import gen_sineembed_for_position


def _get_clones(module, N):
    return nn.ModuleList([copy.deepcopy(module) for _ in range(N)])


class MLP(nn.Module):
    """dab_transformer.py:36-48 (ReLU MLP)."""

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers):
        super().__init__()
        self.num_layers = num_layers
        h = [hidden_dim] * (num_layers - 1)
        self.layers = nn.ModuleList(nn.Linear(n, k) for n, k in zip([input_dim] + h, h + [output_dim]))

    def forward(self, x):
        for i, layer in enumerate(self.layers):
            x = linear(x, layer.weight, layer.bias, _lib.ACT_RELU if i < self.num_layers - 1 else _lib.ACT_NONE)
        return x


class ConvBlock(nn.Module):
    """dab_transformer.py:78-98.  forward takes/returns NCHW like the reference."""

    def __init__(self, dim, drop_path=0):
        super().__init__()
        if drop_path > 0:
            raise NotImplementedError("drop_path > 0 is never used (dab_transformer.py:1017)")
        self.conv1 = nn.Conv2d(dim, dim, kernel_size=(3, 3), padding=1)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.conv2 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.conv3 = nn.Linear(4 * dim, dim)
        self.drop_path = nn.Identity()

    def forward(self, x):
        y = conv_block(x.permute(0, 2, 3, 1).contiguous(), self.conv1.weight, self.conv1.bias, self.norm.weight,
                       self.norm.bias, self.conv2.weight, self.conv2.bias, self.conv3.weight, self.conv3.bias)
        return y.permute(0, 3, 1, 2).contiguous()


class TransformerDecoderLayer(nn.Module):
    """Parameter container + per-layer forward of dab_transformer.py:854-997."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False,
                 keep_query_pos=False, rm_self_attn_decoder=False, num_levels=4):
        super().__init__()
        if rm_self_attn_decoder or keep_query_pos or activation != "relu":
            raise NotImplementedError("only the configuration built by build_transformer is implemented")
        for n in ("sa_qcontent_proj", "sa_qpos_proj", "sa_kcontent_proj", "sa_kpos_proj", "sa_v_proj"):
            setattr(self, n, nn.Linear(d_model, d_model))
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout, vdim=d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.dropout1 = nn.Dropout(dropout)
        self.lvl_w_embed = nn.Linear(d_model, num_levels)
        for n in ("ca_qcontent_proj", "ca_qpos_proj", "ca_kcontent_proj", "ca_kpos_proj", "ca_v_proj", "ca_qpos_sine_proj"):
            setattr(self, n, nn.Linear(d_model, d_model))
        self.cross_attn = MultiheadAttention(d_model * 2, nhead, dropout=dropout, vdim=d_model, query_specific_key=True)
        self.query_specific_key = True
        self.nhead = nhead
        self.rm_self_attn_decoder = rm_self_attn_decoder
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.dropout2 = nn.Dropout(dropout)
        self.dropout3 = nn.Dropout(dropout)
        self.normalize_before = normalize_before
        self.keep_query_pos = keep_query_pos
        self.dim_feedforward = dim_feedforward
        self.dropout_rate = dropout
        self.norm_ = nn.LayerNorm(d_model)

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, query_pos=None, query_sine_embed=None, is_first=False):
        L = lambda m, x, **kw: linear(x, m.weight, m.bias, **kw)
        q = L(self.sa_qpos_proj, query_pos, res=L(self.sa_qcontent_proj, tgt))
        k = L(self.sa_kpos_proj, query_pos, res=L(self.sa_kcontent_proj, tgt))
        v = L(self.sa_v_proj, tgt)
        tgt2 = self.self_attn(q, k, value=v, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)[0]
        tgt = layer_norm(tgt, self.norm1.weight, self.norm1.bias, res=tgt2)
        # level mix (:943-946): tiny softmax + weighted sum as tensor glue, LayerNorm in the kernel
        lvl_w = L(self.lvl_w_embed, tgt).float().softmax(-1).to(tgt.dtype)
        q_memory = torch.einsum("ntl,lhtc->nhtc", lvl_w, memory.to(tgt.dtype))
        q_memory = layer_norm(q_memory, self.norm_.weight, self.norm_.bias)
        q_content = L(self.ca_qcontent_proj, tgt)
        k_content = L(self.ca_kcontent_proj, q_memory)
        v = L(self.ca_v_proj, q_memory)
        nq, bs, C = q_content.shape
        hw = k_content.shape[-3]
        k_pos = L(self.ca_kpos_proj, pos[0:1].to(tgt.dtype)).expand(nq, -1, -1, -1)
        if is_first:
            q = L(self.ca_qpos_proj, query_pos, res=q_content)
            k = k_content + k_pos
        else:
            q, k = q_content, k_content
        H = self.nhead
        qse = L(self.ca_qpos_sine_proj, query_sine_embed.to(tgt.dtype)).view(nq, bs, H, C // H)
        q = torch.cat([q.view(nq, bs, H, C // H), qse], dim=3).view(nq, bs, C * 2)
        k = torch.cat([k.reshape(nq, hw, bs, H, C // H), k_pos.reshape(nq, hw, bs, H, C // H)], dim=4).view(nq, hw, bs, C * 2)
        tgt2 = self.cross_attn(query=q, key=k, value=v, attn_mask=memory_mask, key_padding_mask=memory_key_padding_mask)[0]
        tgt = layer_norm(tgt, self.norm2.weight, self.norm2.bias, res=tgt2)
        tgt_temp = tgt
        tgt2 = L(self.linear2, L(self.linear1, tgt, act=_lib.ACT_RELU))
        tgt = layer_norm(tgt, self.norm3.weight, self.norm3.bias, res=tgt2)
        return tgt, tgt_temp, q_memory


class TransformerClassDecoderLayer(nn.Module):
    """Parameter container + per-layer forward of dab_transformer.py:999-1079."""

    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", num_conv_blocks=3):
        super().__init__()
        self.d_model = d_model
        self.cls_linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout1 = nn.Dropout(dropout)
        self.cls_linear2 = nn.Linear(dim_feedforward, d_model)
        self.dropout2 = nn.Dropout(dropout)
        self.cls_norm = nn.LayerNorm(d_model)
        self.conv_norm = nn.LayerNorm(d_model)
        conv_block_ = ConvBlock(d_model, 0)
        self.conv_blocks = nn.ModuleList([conv_block_ for _ in range(num_conv_blocks)])   # ONE shared block (:1017-1018)
        self.self_attn = MultiheadAttention(d_model, nhead, dropout=dropout, vdim=d_model)
        self.dropout3 = nn.Dropout(dropout)
        self.norm1 = nn.LayerNorm(d_model)
        self.q_proj = nn.Linear(d_model, d_model)       # unused parameter, kept for checkpoint parity (:1026)
        self.k_proj = nn.Conv2d(d_model, d_model, kernel_size=1)
        self.v_proj = nn.Conv2d(d_model, d_model, kernel_size=1)
        self.cls_qpos_sine_proj = nn.Linear(d_model, d_model)
        self.cross_attn = MultiheadAttention(d_model * 2, nhead, dropout=dropout, vdim=d_model)
        self.cls_linear1_ = nn.Linear(d_model, dim_feedforward)
        self.dropout1_ = nn.Dropout(dropout)
        self.cls_linear2_ = nn.Linear(dim_feedforward, d_model)
        self.dropout2_ = nn.Dropout(dropout)
        self.cls_norm_ = nn.LayerNorm(d_model)

    def forward(self, actor_feature, q_memory, pos, query_sine_embed, class_queries, orig_res, num_queries, is_first):
        L = lambda m, x, **kw: linear(x, m.weight, m.bias, **kw)
        dt = actor_feature.dtype
        a2 = L(self.cls_linear2, L(self.cls_linear1, actor_feature, act=_lib.ACT_RELU))
        actor = layer_norm(actor_feature, self.cls_norm.weight, self.cls_norm.bias, res=a2)
        h, w = orig_res
        nq, S, BT, C = q_memory.shape
        N = nq * BT
        enc = q_memory.permute(0, 2, 1, 3).reshape(N, h, w, C).contiguous()                      # NHWC
        feat = layer_norm(enc, self.conv_norm.weight, self.conv_norm.bias,
                          res=actor.reshape(N, 1, 1, C).expand(N, h, w, C))
        blk = self.conv_blocks[0]
        for _ in self.conv_blocks:
            feat = conv_block(feat, blk.conv1.weight, blk.conv1.bias, blk.norm.weight, blk.norm.bias, blk.conv2.weight,
                              blk.conv2.bias, blk.conv3.weight, blk.conv3.bias)
        if is_first:
            query = class_queries.to(dt)[:, None].expand(-1, N, -1).contiguous()
        else:
            query = class_queries.to(dt)
        K = query.shape[0]
        query2 = self.self_attn(query, query, query)[0]
        query = layer_norm(query, self.norm1.weight, self.norm1.bias, res=query2)
        kx = L(self.k_proj, feat.reshape(N, S, C)).permute(1, 0, 2)
        key = torch.cat([kx, pos.to(dt)[:, None].expand(-1, num_queries, -1, -1).flatten(1, 2)], dim=-1).contiguous()
        cqp = L(self.cls_qpos_sine_proj, query_sine_embed.to(dt)).flatten(0, 1)[None].expand(K, -1, -1)
        query_cat = torch.cat([query, cqp], dim=-1).contiguous()
        value = L(self.v_proj, enc.reshape(N, S, C)).permute(1, 0, 2).contiguous()
        out = self.cross_attn(query=query_cat, key=key, value=value)[0]
        cls_output = out.reshape(K, num_queries, -1, self.d_model).permute(1, 2, 0, 3).contiguous()
        c2 = L(self.cls_linear2_, L(self.cls_linear1_, cls_output, act=_lib.ACT_RELU))
        cls_output = layer_norm(cls_output, self.cls_norm_.weight, self.cls_norm_.bias, res=c2)
        next_query = cls_output.permute(2, 0, 1, 3).contiguous().flatten(1, 2)
        return cls_output, next_query


class TransformerDecoder(nn.Module):
    """dab_transformer.py:671-852.  `compute_dtype` selects the bf16 tensor-core path (default) or the fp32 parity path;
    `out_dtype` the dtype of hs / cls_hs (fp32 like the reference by default)."""

    def __init__(self, decoder_layer, cls_decoder_layer, num_layers, norm=None, return_intermediate=False, d_model=256,
                 query_dim=2, keep_query_pos=False, query_scale_type='cond_elewise', modulate_hw_attn=False,
                 bbox_embed_diff_each_layer=False, gradient_checkpointing=False, num_classes=80, temp_len=32):
        super().__init__()
        assert return_intermediate
        if query_scale_type != 'cond_elewise' or not modulate_hw_attn or bbox_embed_diff_each_layer or keep_query_pos \
                or query_dim != 4 or d_model != 256:
            raise NotImplementedError("only the configuration built by build_transformer (dab_transformer.py:1086-1106) "
                                      "with d_model=256 is implemented")
        self.layers = _get_clones(decoder_layer, num_layers)
        self.cls_layers = _get_clones(cls_decoder_layer, num_layers)
        self.num_layers = num_layers
        self.norm = norm
        self.return_intermediate = return_intermediate
        self.query_dim = query_dim
        self.query_scale_type = query_scale_type
        self.query_scale = MLP(d_model, d_model, d_model, 2)
        self.ref_point_head = MLP(query_dim // 2 * d_model, d_model, d_model, 2)
        self.bbox_embed = None          # injected by DETR (models/model.py:100-101)
        self.d_model = d_model
        self.modulate_hw_attn = modulate_hw_attn
        self.bbox_embed_diff_each_layer = bbox_embed_diff_each_layer
        self.ref_anchor_head = MLP(d_model, d_model, 2, 2)
        for layer_id in range(num_layers - 1):
            self.layers[layer_id + 1].ca_qpos_proj = None
        self.cls_norm = nn.LayerNorm(d_model)      # unused parameter (checkpoint parity)
        self.class_queries = nn.Embedding(num_classes, d_model)
        self.temp_len = temp_len
        self.cls_norm2 = nn.LayerNorm(d_model)
        self.gradient_checkpointing = gradient_checkpointing
        self.compute_dtype = torch.bfloat16
        self.out_dtype = torch.float32
        self._engine = None
        self._engine_key = None

    def _get_engine(self, device):
        if self.bbox_embed is None:
            raise RuntimeError("decoder.bbox_embed must be attached (models/model.py:100-101) before forward")
        params = list(self.parameters())
        key = (str(device), self.compute_dtype, self.out_dtype, tuple(p.data_ptr() for p in params))
        ver = tuple(p._version for p in params)
        if self._engine is None or self._engine_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            F = self.layers[0].linear1.out_features
            self._engine = DecoderEngine(sd, nq=None, K=self.class_queries.num_embeddings, layers=self.num_layers, F=F,
                                         dtype=self.compute_dtype, device=device, out_f32=self.out_dtype == torch.float32)
            self._engine_key, self._engine_ver = key, ver
        elif self._engine_ver != ver:        # optimizer step: refresh the packed copies in place, keep workspaces / gradient buffer
            self._engine.repack({k: v for k, v in self.state_dict().items()})
            self._engine_ver = ver
        return self._engine

    def weights_updated(self):
        """Parameters were changed through raw pointers (FlatAdamW): refresh the packed copies at the next forward."""
        self._engine_ver = None

    def forward(self, tgt, memory, tgt_mask=None, memory_mask=None, tgt_key_padding_mask=None,
                memory_key_padding_mask=None, pos=None, refpoints_unsigmoid=None, orig_res=None):
        if tgt_mask is not None or memory_mask is not None or tgt_key_padding_mask is not None:
            raise NotImplementedError("tgt_mask / memory_mask / tgt_key_padding_mask are never passed (dab_transformer.py:395)")
        eng = self._get_engine(tgt.device)
        eng.nq = tgt.shape[0]
        named = [(n, p) for n, p in self.named_parameters()]
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for _, p in named) or tgt.requires_grad or
                                                  memory.requires_grad or refpoints_unsigmoid.requires_grad)
        if needs_grad:
            # training step: native forward that keeps the backward's intermediates + native backward behind autograd.
            # In train() mode nn.Dropout(p) is applied at the residual-branch / FFN-hidden sites (Philox masks, a fresh seed per
            # call); the dropout on attention probabilities (attention.py:402) is not.
            p_drop = float(self.layers[0].dropout1.p) if self.training else 0.0
            self._train_calls = getattr(self, "_train_calls", 0) + 1
            eng.train_dropout = (p_drop, (getattr(self, "dropout_seed", 0x1234) << 24) + self._train_calls)
            from ..functions.decoder_func import DecoderFunction
            hs, cls_hs, refs = DecoderFunction.apply(eng, [n for n, _ in named], memory_key_padding_mask, pos, orig_res, tgt,
                                                     memory, refpoints_unsigmoid, *[p for _, p in named])
            return [hs, cls_hs, refs]
        out = eng.forward(tgt, memory, memory_key_padding_mask, pos, refpoints_unsigmoid, orig_res, heads=False)
        return [out["hs"], out["cls_hs"], out["refs"]]


def build_decoder(num_queries=15, num_classes=80, num_layers=6, dim_feedforward=2048, d_model=256, nhead=8, dropout=0.1,
                  temp_len=16):
    """The decoder exactly as `Transformer.__init__` builds it (dab_transformer.py:138-149) plus the shared
    `bbox_embed` MLP that DETR injects (models/model.py:90,100-101)."""
    layer = TransformerDecoderLayer(d_model, nhead, dim_feedforward, dropout, "relu", False, keep_query_pos=False)
    cls_layer = TransformerClassDecoderLayer(d_model, nhead, dim_feedforward, dropout, "relu", 3)
    dec = TransformerDecoder(layer, cls_layer, num_layers, nn.LayerNorm(d_model), return_intermediate=True,
                             d_model=d_model, query_dim=4, keep_query_pos=False, query_scale_type='cond_elewise',
                             modulate_hw_attn=True, bbox_embed_diff_each_layer=False, num_classes=num_classes,
                             temp_len=temp_len)
    dec.bbox_embed = MLP(d_model, d_model, 4, 3)
    return dec
