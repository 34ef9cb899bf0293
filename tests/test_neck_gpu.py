"""ViT simple-feature-pyramid neck against fixtures from the reference's lateral_convs
(oracle/make_golden_neck.py: the reference's own channel-first LayerNorm class + the module list of backbone_3d_builder.py:139-180)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err, TOL_FP32, TOL_BF16

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["a", "b"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, TOL_FP32), (torch.bfloat16, TOL_BF16)])
def test_vit_neck_matches_reference_lateral_convs(tag, dtype, tol):
    from class_query_vad_b200 import SimpleFeaturePyramid
    from oracle.make_golden_neck import CASES, make_case
    g = load_golden("neck")
    kw = CASES[tag]
    x, sds = make_case(kw)
    dev = torch.device("cuda:0")
    neck = SimpleFeaturePyramid(embed_dim=kw["Cin"])
    sd = {f"lateral_convs.{l}.{k}": torch.from_numpy(v) for l, d in enumerate(sds) for k, v in d.items()}
    neck.load_state_dict(sd, strict=True)              # the reference's parameter names
    neck = neck.to(dev)
    xs = [torch.from_numpy(x).to(dev).to(dtype)] * 4
    out = neck(xs)
    torch.cuda.synchronize()
    for l in range(4):
        ref = g[f"{tag}.{l}"]
        got = out[str(l)].float().cpu().numpy()
        assert got.shape == ref.shape, (l, got.shape, ref.shape)
        assert rel_err(got, ref) < tol, (l, rel_err(got, ref))
    tokens, sh, ls = neck.forward_tokens(xs)
    assert tokens.shape[1] == int(sh.prod(1).sum()) and int(ls[1]) == kw["T"] * 16 * kw["H"] * kw["W"]


def test_vit_neck_batch_independence_at_full_size():
    """Size-independent property at BASELINE configs[0]'s feature size (8 x 14 x 14 x 768 per clip -> 33 320 tokens): clips are
    independent, so the neck of a 2-clip batch equals the two single-clip results bit for bit (tcgen05 path, bf16)."""
    from class_query_vad_b200 import SimpleFeaturePyramid
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    neck = SimpleFeaturePyramid(768).to(dev)
    x = [torch.randn((2, 768, 8, 14, 14), device=dev).bfloat16() for _ in range(4)]
    both, sh, ls = neck.forward_tokens(x)
    assert both.shape == (2, 33320, 256) and torch.isfinite(both.float()).all()
    for b in range(2):
        one, _, _ = neck.forward_tokens([t[b:b + 1].contiguous() for t in x])
        assert torch.equal(one[0], both[b]), b
