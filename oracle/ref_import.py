"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* reference decoder from /root/reference.

Used solely by oracle/make_golden.py (run in the build container, where /root/reference is mounted)
to produce the fixtures under tests/golden/.  Nothing in the product package, bench.py or the `-m gpu`
tests may import this module: /root/reference does not exist on the GPU box.

The reference's `models.detr.dab_transformer` cannot be imported as shipped because of four imports that are
unrelated to the decoder arithmetic:
  * `ops.functions.ms_deform_attn_func` imports the un-built CUDA extension `MultiScaleDeformableAttention`
    (ops/functions/ms_deform_attn_func.py:18)
  * `timm.models.layers.DropPath` (dab_transformer.py:32) -- timm is not installed; ConvBlock is built with
    drop_path=0 so DropPath is never instantiated (dab_transformer.py:86,1017)
  * `VideoMamba.mamba.mamba_ssm.modules.mamba_simple.Mamba` (dab_transformer.py:526) -- un-vendored
  * `selective_scan_cuda_core` & friends (models/detr/common_utils_mbyolo.py:11-28)
They are replaced by inert stand-ins in sys.modules before the import.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("CQVAD_REF", "/root/reference")


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def import_reference():
    """Returns the reference module `models.detr.dab_transformer` (and puts REF_ROOT on sys.path)."""
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch.nn as nn

    def _raise(*a, **k):
        raise RuntimeError("MultiScaleDeformableAttention is a stub (CUDA-only in the reference)")

    if "MultiScaleDeformableAttention" not in sys.modules:
        _mod("MultiScaleDeformableAttention", ms_deform_attn_forward=_raise, ms_deform_attn_backward=_raise)
    if "timm" not in sys.modules:
        class DropPath(nn.Identity):
            def __init__(self, p=0.0):
                super().__init__()
        t = _mod("timm"); t.__path__ = []
        tm = _mod("timm.models"); tm.__path__ = []
        _mod("timm.models.layers", DropPath=DropPath, trunc_normal_=nn.init.trunc_normal_, to_2tuple=lambda x: (x, x))
    for name in ("selective_scan_cuda_core", "selective_scan_cuda_oflex", "selective_scan_cuda_ndstate",
                 "selective_scan_cuda_nrow", "selective_scan_cuda"):
        if name not in sys.modules:
            _mod(name)
    if "VideoMamba" not in sys.modules:
        class Mamba(nn.Module):
            def __init__(self, *a, **k):
                super().__init__()
        chain = "VideoMamba.mamba.mamba_ssm.modules.mamba_simple".split(".")
        for i in range(1, len(chain) + 1):
            m = _mod(".".join(chain[:i])); m.__path__ = []
        sys.modules["VideoMamba.mamba.mamba_ssm.modules.mamba_simple"].Mamba = Mamba
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    return importlib.import_module("models.detr.dab_transformer")
