"""Drop-in for the reference `ops.modules.MSDeformAttn3D` (ops/modules/ms_deform_attn.py:117-203): same constructor,
forward signature, parameter names and initialisation; the four linears run through cqvad_linear, the softmax + sampling
locations through cqvad_msda3d_prepare and the sampling through MSDeformAttnFunction (libcqvad.so).  Stand-alone inference
convenience (the encoder layer's hot path is ONE call, cqvad_deform_encoder_layer_forward): the dtype casts of the offsets /
logits to fp32 and the padding-mask fill are torch tensor glue here.  The constructor and `_reset_parameters` reproduce the
Deformable-DETR initialisation recipe (Apache-2.0, SenseTime; reference NOTICE) because parameter names and initial values are
part of the drop-in contract."""
import math
import warnings

import torch
from torch import nn
from torch.nn.init import xavier_uniform_, constant_

from ..functions.ms_deform_attn_func import MSDeformAttnFunction
from .ops import linear


def _head_dim_is_power_of_two(d_model, n_heads):
    d = d_model // n_heads
    return d > 0 and (d & (d - 1)) == 0


class MSDeformAttn3D(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if not isinstance(d_model, int) or not isinstance(n_heads, int) or n_heads <= 0 or d_model % n_heads != 0:
            raise ValueError(f"MSDeformAttn3D: d_model ({d_model}) has to be a positive multiple of n_heads ({n_heads})")
        if not _head_dim_is_power_of_two(d_model, n_heads):
            warnings.warn("MSDeformAttn3D: a power-of-two head dimension (d_model / n_heads) is what the sampling kernels are laid out for")
        self.im2col_step = 64       # kept for state / attribute compatibility; this library has no im2col batching
        self.d_model, self.n_levels, self.n_heads, self.n_points = d_model, n_levels, n_heads, n_points
        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 3)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)
        self._reset_parameters()

    def _reset_parameters(self):
        """Initial values of ops/modules/ms_deform_attn.py:149-165 (pinned bit for bit by tests/golden/msda_module_init.npz): zero
        offset / attention weights; offset bias = for head h the unit direction (cos, sin) of angle 2 pi (h mod H/2) / (H/2) with
        a time component 1 for the first half of the heads and 0 for the second, normalised by its largest component and scaled
        by (point index + 1), identical for every level; Xavier-uniform value / output projections (in that RNG order)."""
        H, half = self.n_heads, self.n_heads // 2
        angle = torch.arange(half, dtype=torch.float32) * (2.0 * math.pi / half)
        direction = torch.stack([angle.cos().repeat(2), angle.sin().repeat(2),
                                 torch.cat([torch.ones(half), torch.zeros(half)], 0)], -1)
        direction = direction / direction.abs().max(-1, keepdim=True)[0]
        reach = torch.arange(1, self.n_points + 1, dtype=torch.float32).view(1, 1, self.n_points, 1)
        bias = (direction.view(H, 1, 1, 3) * reach).expand(H, self.n_levels, self.n_points, 3)
        with torch.no_grad():
            self.sampling_offsets.weight.zero_()
            self.sampling_offsets.bias = nn.Parameter(bias.reshape(-1).clone())
            self.attention_weights.weight.zero_()
            self.attention_weights.bias.zero_()
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1] * input_spatial_shapes[:, 2]).sum() == Len_in
        value = linear(input_flatten, self.value_proj.weight, self.value_proj.bias)
        if input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        if reference_points.shape[-1] != 3:
            # the 6-vector (reference box) form of ops/modules/ms_deform_attn.py:193-195 is only reached by the two-stage
            # variant, which no shipped configuration enables
            raise ValueError('Last dim of reference_points must be 3, but get {} instead.'.format(reference_points.shape[-1]))
        if self.n_heads != 8:
            raise ValueError("libcqvad MSDeformAttn3D: 8 heads (every shipped configuration)")
        offs = linear(query, self.sampling_offsets.weight, self.sampling_offsets.bias)
        attw = linear(query, self.attention_weights.weight, self.attention_weights.bias)
        # softmax over the L*P logits + sampling locations (reference quirk kept: normaliser stacked as (T_l, W_l, H_l) against
        # (x, y, t) offsets, ms_deform_attn.py:190) in one kernel: cqvad_msda3d_prepare
        from .. import _lib
        rows = N * Len_q
        off32, lg32 = offs.float().contiguous(), attw.float().contiguous()
        ref32 = reference_points.float().contiguous()
        sampling_locations = torch.empty((N, Len_q, self.n_heads, self.n_levels, self.n_points, 3), dtype=torch.float32,
                                         device=query.device)
        attw = torch.empty((N, Len_q, self.n_heads, self.n_levels, self.n_points), dtype=torch.float32, device=query.device)
        p = _lib.ptr
        _lib.check(_lib.lib().cqvad_msda3d_prepare(p(off32), p(lg32), p(ref32), p(input_spatial_shapes.to(torch.int64).contiguous()),
                                                  p(sampling_locations), p(attw), rows, self.n_levels, self.n_points,
                                                  _lib.stream_ptr()))
        output = MSDeformAttnFunction.apply(value.contiguous(), input_spatial_shapes, input_level_start_index,
                                            sampling_locations.contiguous(), attw.contiguous(), self.im2col_step)
        return linear(output, self.output_proj.weight, self.output_proj.bias)
