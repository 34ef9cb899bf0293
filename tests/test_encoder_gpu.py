"""GPU parity of the deformable encoder layer through the C ABI
(cqvad_deform_encoder_layer_forward) and through the drop-in modules, against fixtures produced by the UNMODIFIED reference
DeformableTransformerEncoderLayer / MSDeformAttn3D module / get_reference_points (oracle/make_golden_encoder.py).
Tolerances: BASELINE.json's rel 1e-3 in fp32, 2e-2 in bf16 (rel = |a-b|_inf / |b|_inf)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err, TOL_FP32, TOL_BF16
from oracle import synth
from test_oracle_golden import enc_case, ENC_CASES

pytestmark = pytest.mark.gpu


def _run(g, dtype, want_attn=True):
    from class_query_vad_b200 import pack_encoder_layer_weights, encoder_layer_forward
    W, inp, shapes, masked = enc_case(g)
    B, F_, P, seed, _ = (int(v) for v in g["meta"])
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    packed = pack_encoder_layer_weights({k: torch.from_numpy(v) for k, v in W.items()}, dtype, dev)
    out = encoder_layer_forward(packed, t(inp["src"]).to(dtype), t(inp["pos"]).to(dtype), t(g["reference_points"]), t(g["shapes"]),
                                t(g["level_start"]), t(inp["mask"]) if masked else None, P, F_, want_attn=want_attn)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize("name", ENC_CASES)
def test_encoder_layer_fp32_matches_reference(name):
    g = load_golden(name)
    out, attn_out = _run(g, torch.float32)
    assert rel_err(attn_out.cpu().numpy(), g["attn_out"]) < TOL_FP32
    assert rel_err(out.cpu().numpy(), g["out"]) < TOL_FP32


@pytest.mark.parametrize("name", ENC_CASES)
def test_encoder_layer_bf16_matches_reference(name):
    g = load_golden(name)
    out, attn_out = _run(g, torch.bfloat16)
    assert rel_err(attn_out.float().cpu().numpy(), g["attn_out"]) < TOL_BF16
    assert rel_err(out.float().cpu().numpy(), g["out"]) < TOL_BF16


def test_encoder_module_drop_in_state_dict_and_forward():
    """The module loads the reference-named state_dict strictly, computes the reference points itself and stacks layers."""
    from class_query_vad_b200 import DeformableTransformerEncoderLayer, DeformableTransformerEncoder
    g = load_golden("enc_small_masked")
    W, inp, shapes, masked = enc_case(g)
    B, F_, P, seed, _ = (int(v) for v in g["meta"])
    dev = torch.device("cuda:0")
    layer = DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, n_levels=len(shapes), n_heads=8, n_points=P)
    layer.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
    enc = DeformableTransformerEncoder(layer, 1).to(dev).eval()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = enc(t(inp["src"]), t(g["shapes"]), t(g["level_start"]), t(inp["valid_ratios"]), pos=t(inp["pos"]), padding_mask=t(inp["mask"]))
    assert rel_err(out.cpu().numpy(), g["out"]) < TOL_FP32
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        DeformableTransformerEncoder(layer, 1).eval()(torch.from_numpy(inp["src"]), torch.from_numpy(g["shapes"]),
                                                      torch.from_numpy(g["level_start"]), torch.from_numpy(inp["valid_ratios"]),
                                                      pos=torch.from_numpy(inp["pos"]))


def test_encoder_layer_full_pyramid_batch_independence():
    """Full ViT-B/224 pyramid (Len = 33 320 tokens per clip): clips are independent, so a 2-clip batch
    must reproduce each clip run alone bit for bit (size-independent property at BASELINE size), in bf16 on the tcgen05 path."""
    from class_query_vad_b200 import pack_encoder_layer_weights, encoder_layer_forward
    from oracle import encoder_np
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    Wn = synth.make_encoder_layer_weights(2048, 4, 8, seed=5)
    inp = synth.make_encoder_inputs(2, shapes, seed=5)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    packed = pack_encoder_layer_weights({k: torch.from_numpy(v) for k, v in Wn.items()}, torch.bfloat16, dev)
    refp = t(encoder_np.reference_points(shapes, inp["valid_ratios"]))
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    src, pos = t(inp["src"]).bfloat16(), t(inp["pos"]).bfloat16()
    both = encoder_layer_forward(packed, src, pos, refp, sh, ls, None, 8, 2048)
    one = encoder_layer_forward(packed, src[1:2], pos[1:2], refp[1:2], sh, ls, None, 8, 2048)
    torch.cuda.synchronize()
    assert src.shape[1] == 33320
    assert torch.isfinite(both.float()).all()
    assert torch.equal(both[1:2], one)


def test_encoder_layer_fused_sampling_variant_matches_reference():
    """CQVAD_ENC_FUSED=1 (softmax + sampling locations inside the sampling kernel; read once per process, hence a subprocess)."""
    import os, subprocess, sys
    code = ("import sys, torch; sys.path.insert(0, 'tests'); "
            "from helpers import load_golden, rel_err; from test_encoder_gpu import _run; "
            "g = load_golden('enc_small_masked'); "
            "o32, a32 = _run(g, torch.float32); o16, a16 = _run(g, torch.bfloat16); "
            "assert rel_err(o32.cpu().numpy(), g['out']) < 1e-3 and rel_err(a32.cpu().numpy(), g['attn_out']) < 1e-3; "
            "assert rel_err(o16.float().cpu().numpy(), g['out']) < 2e-2; print('fused ok')")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=dict(os.environ, CQVAD_ENC_FUSED="1"), capture_output=True,
                       text=True, timeout=300)
    assert r.returncode == 0 and "fused ok" in r.stdout, r.stderr[-2000:]
