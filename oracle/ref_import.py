"""TEST / BASELINE INFRASTRUCTURE ONLY -- imports the *unmodified* reference decoder.

Root of the reference tree: $CQVAD_REF, else the git-ignored install `baseline/_ref/` (oracle/install_ref.py: a byte-for-byte
copy that travels to the GPU box), else /root/reference (build container only).  Users: the fixture generators
oracle/make_golden*.py, `bench.py --impl reference` / its `cpu_baseline` and `reference_gpu_eager` legs (the reference timed
beside the product, never as the product), and tests/test_msda_ref_gpu.py.  Nothing in the product package imports this module.

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

_HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    for c in (os.environ.get("CQVAD_REF"), os.path.join(_HERE, "baseline", "_ref"), "/root/reference"):
        if c and os.path.isdir(os.path.join(c, "models", "detr")):
            return c
    return os.environ.get("CQVAD_REF", "/root/reference")


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "models", "detr"))


def import_reference_msda():
    """The reference's own compiled CUDA extension (ops/src built by oracle/install_ref.py with the value.scalar_type() shim);
    returns the module with ms_deform_attn_forward / ms_deform_attn_backward, or None when it was not built."""
    import importlib
    import torch  # noqa: F401  (libtorch must be loaded first)
    so = os.path.join(REF_ROOT, "MultiScaleDeformableAttention.so")
    if not os.path.exists(so):
        return None
    m = sys.modules.get("MultiScaleDeformableAttention")
    if m is not None and getattr(m, "__file__", None):
        return m
    sys.modules.pop("MultiScaleDeformableAttention", None)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("MultiScaleDeformableAttention")


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

    if "MultiScaleDeformableAttention" not in sys.modules and (
            os.environ.get("CQVAD_REF_STUB_MSDA", "1") == "1" or import_reference_msda() is None):
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
