"""Drop-in contract of MSDeformAttn3D's constructor: the initial parameters are the reference's own, bit for bit
(ops/modules/ms_deform_attn.py:117-165; fixture from the unmodified reference module, oracle/make_golden_msda_init.py)."""
import os
import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_initial_parameters_match_the_reference_module():
    from class_query_vad_b200.modules.ms_deform_attn import MSDeformAttn3D
    g = np.load(os.path.join(GOLD, "msda_module_init.npz"))
    torch.manual_seed(0)
    m = MSDeformAttn3D(d_model=256, n_levels=4, n_heads=8, n_points=8)
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    assert list(sd) == ["sampling_offsets.weight", "sampling_offsets.bias", "attention_weights.weight", "attention_weights.bias",
                        "value_proj.weight", "value_proj.bias", "output_proj.weight", "output_proj.bias"]
    assert np.array_equal(sd["sampling_offsets.bias"], g["sampling_offsets_bias"])
    assert np.array_equal(sd["value_proj.weight"][:4, :8], g["value_proj_corner"])          # same RNG order as the reference
    assert np.array_equal(sd["output_proj.weight"][:4, :8], g["output_proj_corner"])
    assert sd["value_proj.weight"].astype(np.float64).sum() == float(g["value_proj_sum"])
    assert sd["output_proj.weight"].astype(np.float64).sum() == float(g["output_proj_sum"])
    for k in ("sampling_offsets.weight", "attention_weights.weight", "attention_weights.bias", "value_proj.bias", "output_proj.bias"):
        assert not sd[k].any()
    assert float(g["zero_max"]) == 0.0


def test_constructor_rejects_bad_geometry():
    from class_query_vad_b200.modules.ms_deform_attn import MSDeformAttn3D
    with pytest.raises(ValueError):
        MSDeformAttn3D(d_model=250, n_heads=8)
    with pytest.warns(UserWarning):
        MSDeformAttn3D(d_model=240, n_levels=1, n_heads=8, n_points=1)      # head dimension 30
