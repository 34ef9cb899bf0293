"""DeformableTransformerEncoder.get_reference_points (broadcast formulation) against the reference's own function
(models/detr/dab_transformer.py:433-449) through the fixtures the unmodified reference produced (tests/golden/enc_*.npz):
bit for bit, since these values enter the bit-exact sampling-index contract of the MSDA op."""
import numpy as np
import pytest
import torch

from helpers import load_golden
from test_oracle_golden import enc_case, ENC_CASES


@pytest.mark.parametrize("name", ENC_CASES + ["enc_grad_tiny", "enc_grad_small_masked"])
def test_reference_points_bit_exact(name):
    from class_query_vad_b200.modules.encoder import DeformableTransformerEncoder
    g = load_golden(name)
    _, inp, shapes, _ = enc_case(g)
    vr = torch.from_numpy(np.ascontiguousarray(inp["valid_ratios"]))
    got = DeformableTransformerEncoder.get_reference_points(torch.tensor(shapes, dtype=torch.int64), vr, device=torch.device("cpu"))
    assert got.dtype == torch.float32 and tuple(got.shape) == tuple(g["reference_points"].shape)
    assert np.array_equal(got.numpy(), g["reference_points"])
