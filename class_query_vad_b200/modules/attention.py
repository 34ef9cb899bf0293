"""Drop-in for the reference projection-free `MultiheadAttention` (models/detr/attention.py:61-187): same constructor
arguments, the single `out_proj` parameter, the three call modes and the (output, weights) return convention.
The head-averaged attention map that the reference computes by default and every caller discards
(attention.py:417-420) is not produced: the second return value is always None."""
import torch
from torch import nn
from torch.nn.init import constant_

from .ops import linear, mha_core


class MultiheadAttention(nn.Module):
    def __init__(self, embed_dim, num_heads, dropout=0., bias=True, add_bias_kv=False, add_zero_attn=False, kdim=None,
                 vdim=None, query_specific_key=False, stop_middle=False):
        super().__init__()
        if add_bias_kv or add_zero_attn or stop_middle:
            raise NotImplementedError("add_bias_kv / add_zero_attn / stop_middle are never used on the decoder path")
        self.embed_dim = embed_dim
        self.kdim = kdim if kdim is not None else embed_dim
        self.vdim = vdim if vdim is not None else embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == self.embed_dim, "embed_dim must be divisible by num_heads"
        self.out_proj = nn.Linear(self.vdim, self.vdim)
        self.query_specific_key = query_specific_key
        constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, key_padding_mask=None, need_weights=True, attn_mask=None):
        if attn_mask is not None:
            raise NotImplementedError("attn_mask is always None on the decoder path (dab_transformer.py:933,987)")
        if self.training and self.dropout > 0:
            raise NotImplementedError("training-mode attention dropout is not implemented; call .eval()")
        o = mha_core(query, key, value, self.num_heads, key_padding_mask, self.query_specific_key)
        return linear(o, self.out_proj.weight, self.out_proj.bias), None
