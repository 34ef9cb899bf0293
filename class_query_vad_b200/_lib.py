"""ctypes binding of libcqvad.so (include/cqvad.h).  There is NO fallback: if the library is missing the import of any
op fails loudly -- build it with `python -c "import __graft_entry__ as g; g.build()"` (or `make -C class_query_vad_b200/csrc`)."""
import ctypes
import os
from ctypes import c_int, c_long, c_size_t, c_void_p, c_float, c_char_p, c_uint64, c_uint32, c_uint, POINTER, Structure

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcqvad.so")

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
DEC_SKIP_CLS_HS = 1
DEC_FP32_CLS_STREAM = 2


class DecoderDesc(Structure):
    _fields_ = [("dtype", c_int), ("BT", c_int), ("nq", c_int), ("h", c_int), ("w", c_int), ("K", c_int), ("F", c_int),
                ("layers", c_int), ("out_f32", c_int), ("flags", c_int), ("dropout_p", c_float), ("seed_lo", c_uint), ("seed_hi", c_uint)]


class CriterionCfg(Structure):
    _fields_ = [(n, c_float) for n in ("cost_class", "cost_bbox", "cost_giou", "w_ce", "w_bbox", "w_giou", "w_ce_b", "pos_weight",
                                       "eos_coef", "focal_alpha", "focal_gamma", "label_smoothing")]


class CqvadError(RuntimeError):
    pass


# every symbol include/cqvad.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "cqvad_version": (c_int, []),
    "cqvad_last_error": (c_char_p, []),
    "cqvad_msda3d_forward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 7 + [c_void_p]),
    "cqvad_msda3d_backward": (c_int, [c_int] + [c_void_p] * 9 + [c_int] * 7 + [c_void_p]),
    "cqvad_msda3d_indices": (c_int, [c_void_p] * 6 + [c_int] * 5 + [c_void_p]),
    "cqvad_layernorm": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_int, c_long, c_int, c_void_p]),
    "cqvad_linear_gelu_train": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p]),
    "cqvad_linear_dgrad_act": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_long, c_int, c_int, c_void_p]),
    "cqvad_deform_encoder_layer_train_workspace_bytes": (c_size_t, [c_int, c_int, c_long, c_int, c_int, c_int]),
    "cqvad_deform_encoder_layer_train_forward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                          c_void_p, c_void_p, c_size_t, c_int, c_long, c_int, c_int, c_int, c_float, c_uint64,
                                                          c_void_p]),
    "cqvad_deform_encoder_layer_backward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                     c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_long, c_int, c_int, c_int, c_float, c_uint64,
                                                     c_void_p]),
    "cqvad_input_proj_3x3s2_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "cqvad_input_proj_3x3s2_gn": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                           c_int, c_int, c_int, c_int, c_int, c_long, c_long, c_void_p]),
    "cqvad_input_proj_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_long]),
    "cqvad_input_proj_1x1_gn": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_size_t,
                                         c_int, c_int, c_long, c_long, c_long, c_void_p]),
    "cqvad_msda3d_prepare": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_void_p]),
    "cqvad_level_to_tokens": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_long, c_long, c_long, c_void_p]),
    "cqvad_encoder_to_decoder_memory": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_int, c_int, c_int,
                                                 c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cqvad_vit_neck_workspace_bytes": (c_size_t, [c_int] * 7),
    "cqvad_vit_neck_level": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_long, c_long, c_void_p, c_size_t] + [c_int] * 5 + [c_void_p]),
    "cqvad_level_to_tokens_backward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_long, c_long, c_long, c_void_p]),
    "cqvad_encoder_to_decoder_memory_backward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_long, c_int, c_int, c_int,
                                                          c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cqvad_deform_encoder_layer_num_weights": (c_int, []),
    "cqvad_deform_encoder_layer_workspace_bytes": (c_size_t, [c_int, c_int, c_long, c_int, c_int, c_int]),
    "cqvad_deform_encoder_layer_forward": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                    c_void_p, c_void_p, c_size_t, c_int, c_long, c_int, c_int, c_int, c_void_p]),
    "cqvad_linear": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_void_p]),
    "cqvad_mlp": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_float,
                          c_void_p, c_void_p, c_long, c_int, c_void_p]),
    "cqvad_convblock_workspace_bytes": (c_size_t, [c_int, c_long, c_int, c_int]),
    "cqvad_convblock_forward": (c_int, [c_int] + [c_void_p] * 10 + [c_long, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "cqvad_mha_core": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 6 + [c_void_p]),
    "cqvad_posenc3d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cqvad_sine_embed": (c_int, [c_void_p, c_void_p, c_long, c_void_p]),
    "cqvad_decoder_num_weights": (c_int, [c_int]),
    "cqvad_decoder_weight_name": (c_char_p, [c_int, c_int]),
    "cqvad_decoder_weight_kind": (c_int, [c_int, c_int]),
    "cqvad_decoder_workspace_bytes": (c_size_t, [POINTER(DecoderDesc)]),
    "cqvad_decoder_forward": (c_int, [POINTER(DecoderDesc), POINTER(c_void_p)] + [c_void_p] * 11 + [c_void_p, c_size_t, c_void_p]),
    "cqvad_decoder_train_workspace_bytes": (c_size_t, [POINTER(DecoderDesc)]),
    "cqvad_decoder_train_forward": (c_int, [POINTER(DecoderDesc), POINTER(c_void_p)] + [c_void_p] * 8 + [c_void_p, c_size_t, c_void_p]),
    "cqvad_decoder_backward": (c_int, [POINTER(DecoderDesc), POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p,
                                       POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "cqvad_wgrad_workspace_bytes": (c_size_t, []),
    "cqvad_linear_wgrad": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_long, c_int, c_int, c_int, c_int, c_void_p,
                                   c_size_t, c_void_p]),
    "cqvad_criterion_ava_workspace_bytes": (c_size_t, [c_int]),
    "cqvad_criterion_ava": (c_int, [POINTER(CriterionCfg)] + [c_void_p] * 6 + [c_int] * 4 + [c_void_p] * 6 + [c_size_t, c_void_p]),
    "cqvad_postprocess_ava": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "cqvad_postprocess_ucf": (c_int, [c_void_p] * 5 + [c_int] * 3 + [c_void_p]),
    "cqvad_dropout": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_long, c_float, c_uint64, c_uint32, c_void_p]),
    "cqvad_heads_train_workspace_bytes": (c_size_t, [c_long]),
    "cqvad_heads_train_forward": (c_int, [c_void_p] * 4 + [c_long, c_int, c_float, c_uint64] + [c_void_p] * 4 + [c_size_t, c_void_p]),
    "cqvad_heads_train_backward": (c_int, [c_void_p] * 6 + [c_long, c_int, c_float, c_uint64] + [c_void_p] * 5 + [c_size_t, c_void_p]),
    "cqvad_decoder_backward_layer_events": (c_int, [c_void_p, c_int]),
    "cqvad_adamw_workspace_bytes": (c_size_t, []),
    "cqvad_adamw_clip_step": (c_int, [c_void_p] * 5 + [c_long] + [c_float] * 5 + [c_long, c_float, c_float, c_int, c_void_p, c_void_p,
                                                                                c_size_t, c_void_p]),
    "cqvad_last_launch_count": (c_long, []),
    "cqvad_profile_enable": (None, [c_int]),
    "cqvad_profile_timeline": (c_long, [c_void_p, c_void_p, c_void_p, c_void_p, c_long]),
    "cqvad_profile_num_classes": (c_int, []),
    "cqvad_profile_class_name": (c_char_p, [c_int]),
    "cqvad_profile_read": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(c_long), POINTER(c_long)]),
    "cqvad_profile_read_work": (c_int, [c_int, POINTER(ctypes.c_double), POINTER(ctypes.c_double)]),
    "cqvad_debug_force_simt": (None, [c_int]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                              "(python -c 'import __graft_entry__ as g; g.build()'). There is no CPU/PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise CqvadError(f"libcqvad error {rc}: {lib().cqvad_last_error().decode()}")


def ptr(t):
    """device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def dtype_id(torch_dtype):
    import torch
    if torch_dtype == torch.float32:
        return F32
    if torch_dtype == torch.bfloat16:
        return BF16
    raise CqvadError(f"unsupported dtype {torch_dtype}: libcqvad computes in float32 or bfloat16")


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("Not implemented on the CPU")  # same message as ops/src/cpu/ms_deform_attn_cpu.cpp:26


# ---- level-shape tensors: built once per (shapes, device), with their host copies kept beside them so that the forward path
# never reads a shape back from the device (a .tolist() / .item() there is a full stream synchronisation per call)
_SHAPE_CACHE = {}
_SHAPE_HOST = {}


def shape_tensors(shapes, device):
    """shapes: sequence of (T, H, W).  Returns (shapes [L,3] int64, level_start [L] int64) on `device` (cached)."""
    import torch
    key = (tuple(tuple(int(v) for v in s) for s in shapes), str(device))
    hit = _SHAPE_CACHE.get(key)
    if hit is None:
        sh = torch.tensor(key[0], dtype=torch.int64, device=device)
        ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
        hit = (sh, ls)
        _SHAPE_CACHE[key] = hit
        _SHAPE_HOST[(sh.data_ptr(), str(device))] = [list(s) for s in key[0]]
    return hit


def host_shapes(sh):
    """[[T, H, W], ...] of a level-shape tensor: from the cache when it came from shape_tensors(), else one device read."""
    if not hasattr(sh, "data_ptr"):
        return [list(s) for s in sh]
    hit = _SHAPE_HOST.get((sh.data_ptr(), str(sh.device)))
    return hit if hit is not None else sh.tolist()


_WARNED_NO_GRAD = set()


def warn_if_grad(what, *tensors):
    """The building-block wrappers (modules/ops.py, MSDeformAttn3D.forward, the stand-alone per-layer forwards) launch kernels through
    raw pointers and record NO autograd graph.  Called with gradients enabled on tensors that require them they would silently cut
    the graph, so they say so once per entry point; training goes through TransformerDecoder / Transformer / the encoder layer
    modules, whose forwards are autograd Functions."""
    import torch
    if what in _WARNED_NO_GRAD or not torch.is_grad_enabled():
        return
    if any(t is not None and getattr(t, "requires_grad", False) for t in tensors):
        import warnings
        _WARNED_NO_GRAD.add(what)
        warnings.warn(f"class_query_vad_b200.{what} is an inference building block: no autograd graph is recorded, gradients will not "
                      "flow through it (use torch.no_grad(), or the module-level training paths)", stacklevel=3)
