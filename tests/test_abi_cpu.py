"""CPU-side checks of the C ABI: the library loads, exports every symbol include/cqvad.h declares, and its argument
validation / weight-table enumeration work without a GPU (no compute calls)."""
import ctypes
import os
import re
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from class_query_vad_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.lib()


def test_header_symbols_exported(lib):
    from class_query_vad_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cqvad.h")).read()
    declared = set(re.findall(r"\b(cqvad_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"cqvad_decoder_desc"}
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/cqvad.h but not exported"
    assert declared == set(_lib.SYMBOLS), (declared ^ set(_lib.SYMBOLS))


def test_version_and_error_string(lib):
    assert lib.cqvad_version() == 1
    rc = lib.cqvad_layernorm(0, None, None, None, None, 1e-5, None, 0, 4, 256, None)
    assert rc == -1 and b"layernorm" in lib.cqvad_last_error()
    rc = lib.cqvad_msda3d_forward(0, None, None, None, None, None, None, 1, 1, 1, 200, 1, 1, 1, None)
    assert rc == -2  # D > 128 unsupported


def test_weight_table_matches_reference_names(lib):
    from oracle import synth
    for layers in (1, 3, 6):
        n = lib.cqvad_decoder_num_weights(layers)
        names = [lib.cqvad_decoder_weight_name(i, layers).decode() for i in range(n)]
        assert len(set(names)) == n
        spec = dict(synth.decoder_param_spec(80, layers, 2048))
        for i, nm in enumerate(names):
            if "ca_qpos_proj" in nm and not nm.startswith("layers.0."):
                assert nm not in spec
                continue
            if ".__" in nm:      # synthesised (stacked) weights built by pack_decoder_weights
                assert nm not in spec
                continue
            assert nm in spec, nm
            kind = lib.cqvad_decoder_weight_kind(i, layers)
            assert kind in (0, 1)
            if nm.endswith(".bias") or "norm" in nm.split(".")[-2]:
                assert kind == 1
        # everything the reference forward actually uses is in the table (unused: cls_norm, q_proj; aliases 1,2)
        unused = [k for k in spec if k not in names]
        assert all(("q_proj" in k) or k.startswith("cls_norm.") for k in unused), unused
        assert lib.cqvad_decoder_weight_name(n, layers) is None


def test_workspace_query(lib):
    from class_query_vad_b200._lib import DecoderDesc
    d = DecoderDesc(1, 32, 15, 14, 14, 80, 2048, 6, 1, 0)
    b = lib.cqvad_decoder_workspace_bytes(ctypes.byref(d))
    assert 100e6 < b < 4e9
    d2 = DecoderDesc(0, 32, 15, 14, 14, 80, 2048, 6, 1, 0)
    assert lib.cqvad_decoder_workspace_bytes(ctypes.byref(d2)) > b
    bad = DecoderDesc(1, 32, 15, 14, 200, 80, 2048, 6, 1, 0)
    assert lib.cqvad_decoder_workspace_bytes(ctypes.byref(bad)) == 0


def test_ops_fail_loudly_on_cpu_tensors():
    import torch
    from class_query_vad_b200 import MSDeformAttnFunction
    v = torch.zeros(1, 4, 1, 4)
    shapes = torch.tensor([[1, 2, 2]], dtype=torch.int64)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        MSDeformAttnFunction.apply(v, shapes, torch.zeros(1, dtype=torch.int64), torch.zeros(1, 2, 1, 1, 1, 3),
                                   torch.ones(1, 2, 1, 1, 1), 64)


def test_drop_in_state_dict_names_match_reference():
    """load_state_dict compatibility: parameter names/shapes of the drop-in decoder == synth spec (= reference)."""
    from class_query_vad_b200 import build_decoder
    from oracle import synth
    dec = build_decoder(num_queries=15, num_classes=80, num_layers=2, dim_feedforward=2048)
    sd = dec.state_dict()
    spec = {k: v for k, v in synth.decoder_param_spec(80, 2, 2048) if not k.startswith("heads.")}
    for k, shape in spec.items():
        assert k in sd and tuple(sd[k].shape) == tuple(shape), k
    extra = [k for k in sd if k not in spec and ".conv_blocks.1." not in k and ".conv_blocks.2." not in k]
    assert not extra, extra
