"""Drop-in for the reference `ops.functions.MSDeformAttnFunction` (ops/functions/ms_deform_attn_func.py:21-45).

Same `apply(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights, im2col_step)`
signature and saved tensors; the pybind module `MultiScaleDeformableAttention` (ops/src/vision.cpp:13-16) is replaced
by cqvad_msda3d_forward / cqvad_msda3d_backward of libcqvad.so.

Differences, all deliberate:
  * `im2col_step` is accepted and ignored (one launch covers the batch; no `batch % im2col_step` restriction,
    ms_deform_attn_cuda_t.cu:50-52).
  * the backward is the mathematical gradient of the forward; the reference backward kernel is not.
  * no silent fp16 retry on exceptions (ms_deform_attn_func.py:25-33): errors are raised.
  * value may be float32 or bfloat16; sampling locations / attention weights are consumed in float32.
"""
import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .. import _lib


def _dims(value, loc):
    N, Len, M, D = value.shape
    _, Lq, _, L, P, three = loc.shape
    if three != 3:
        raise ValueError("sampling_locations must have a last dimension of 3 (x, y, t)")
    return N, Len, M, D, L, Lq, P


def _prep(value, shapes, lsi, loc, attn):
    _lib.require_cuda(value, shapes, lsi, loc, attn)
    if shapes.dtype != torch.int64 or lsi.dtype != torch.int64:
        raise TypeError("value_spatial_shapes and value_level_start_index must be int64 (T,H,W)")
    for name, t in (("value", value), ("spatial_shapes", shapes), ("level_start_index", lsi),
                    ("sampling_loc", loc), ("attn_weight", attn)):
        if not t.is_contiguous():
            raise RuntimeError(f"{name} tensor has to be contiguous")  # ms_deform_attn_cuda_t.cu:28-32
    return loc.float().contiguous(), attn.float().contiguous()


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
                im2col_step=64):
        ctx.im2col_step = im2col_step
        loc32, attn32 = _prep(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        N, Len, M, D, L, Lq, P = _dims(value, loc32)
        out = torch.empty((N, Lq, M * D), dtype=value.dtype, device=value.device)
        p = _lib.ptr
        _lib.check(_lib.lib().cqvad_msda3d_forward(_lib.dtype_id(value.dtype), p(value), p(value_spatial_shapes),
                                                   p(value_level_start_index), p(loc32), p(attn32), p(out),
                                                   N, Len, M, D, L, Lq, P, _lib.stream_ptr()))
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights)
        return out

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        value, shapes, lsi, loc, attn = ctx.saved_tensors
        loc32, attn32 = loc.float().contiguous(), attn.float().contiguous()
        N, Len, M, D, L, Lq, P = _dims(value, loc32)
        go = grad_output.to(value.dtype).contiguous()
        g_value = torch.zeros(value.shape, dtype=torch.float32, device=value.device)
        g_loc = torch.empty(loc.shape, dtype=torch.float32, device=value.device)
        g_attn = torch.empty(attn.shape, dtype=torch.float32, device=value.device)
        p = _lib.ptr
        _lib.check(_lib.lib().cqvad_msda3d_backward(_lib.dtype_id(value.dtype), p(value), p(shapes), p(lsi), p(loc32),
                                                    p(attn32), p(go), p(g_value), p(g_loc), p(g_attn),
                                                    N, Len, M, D, L, Lq, P, _lib.stream_ptr()))
        return g_value.to(value.dtype), None, None, g_loc.to(loc.dtype), g_attn.to(attn.dtype), None


def ms_deform_attn_indices(value_spatial_shapes, sampling_locations):
    """Integer part of the sampling: (t_low, h_low, w_low) int32 and the 8-bit corner-validity mask per
    (n, q, m, l, p) -- the 'bit-exact sampling indices' contract (cuh:38-60, 424-428)."""
    _lib.require_cuda(value_spatial_shapes, sampling_locations)
    loc = sampling_locations.float().contiguous()
    N, Lq, M, L, P, _ = loc.shape
    shp = (N, Lq, M, L, P)
    tl = torch.empty(shp, dtype=torch.int32, device=loc.device)
    hl, wl = torch.empty_like(tl), torch.empty_like(tl)
    mask = torch.empty(shp, dtype=torch.uint8, device=loc.device)
    p = _lib.ptr
    _lib.check(_lib.lib().cqvad_msda3d_indices(p(value_spatial_shapes.contiguous()), p(loc), p(tl), p(hl), p(wl), p(mask),
                                               N, Lq, M, L, P, _lib.stream_ptr()))
    return tl, hl, wl, mask
