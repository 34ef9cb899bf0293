"""Pins the CPU oracle (oracle/*.py) against fixtures produced by the reference itself (oracle/make_golden.py)."""
import numpy as np
import pytest
from helpers import load_golden, case_from_meta, case_from_golden, rel_err
from oracle import decoder_np, posenc_np, msda_np, synth

FAST = ["dec_tiny", "dec_tiny_masked", "dec_small_masked", "dec_ucf_like", "dec_jhmdb_like", "dec_ava_csn_b1_l2"]


def _check_decoder(name, tol=2e-5):
    g = load_golden(name)
    cfg, B, W, inp = case_from_meta(g["meta"])
    taps = {}
    hs, cls_hs, refs = decoder_np.decoder_forward(W, inp["tgt"], inp["memory"], inp["mask"], inp["pos"],
                                                  inp["refpoints_unsigmoid"], inp["orig_res"], cfg["layers"], taps=taps)
    logits, boxes, logits_b = decoder_np.detr_heads(W, hs, cls_hs, refs)
    assert rel_err(hs, g["hs"]) < tol
    assert rel_err(refs, g["refs"]) < tol
    assert rel_err(logits, g["pred_logits"]) < tol
    assert rel_err(boxes, g["pred_boxes"]) < tol
    assert rel_err(logits_b, g["pred_logits_b"]) < tol
    if "cls_hs" in g:
        assert rel_err(cls_hs, g["cls_hs"]) < tol
    else:
        assert rel_err(cls_hs[:, :, ::4, ::7, ::5], g["cls_hs_sub"]) < tol
        assert rel_err((cls_hs.astype(np.float64) ** 2).mean(-1), g["cls_hs_sq"]) < tol
    for k in g:
        if k.startswith("l") and "." in k:
            assert rel_err(taps[k], g[k]) < tol, k


@pytest.mark.parametrize("name", FAST)
def test_decoder_oracle_matches_reference(name):
    _check_decoder(name)


def test_decoder_oracle_matches_reference_ava_vitb_full():
    # BASELINE.json configs[0]: AVA22_ViT-B decoder forward, batch 1, 6 layers, K=80, S=196
    _check_decoder("dec_ava_vitb_b1", tol=5e-5)


def test_posenc_oracle():
    g = load_golden("posenc")
    pos = posenc_np.position_embedding_sine_3d(g["mask"])
    assert rel_err(pos, g["pos"]) < 1e-5
    pos2 = posenc_np.position_embedding_sine_3d(np.zeros((1, 8, 14, 14), dtype=bool))
    assert rel_err(pos2[:, :, ::3, ::5, ::4], g["pos_vit14"]) < 1e-5
    assert rel_err(posenc_np.gen_sineembed_for_position(g["ref_in"]), g["sine"]) < 1e-5


def test_attention_oracle():
    g = load_golden("attention")
    o = decoder_np.mha_standard(g["a_q"], g["a_k"], g["a_v"], 8, g["a_wo"], g["a_bo"], key_padding_mask=g["a_kpm"])
    assert rel_err(o, g["a_out"]) < 1e-5
    o = decoder_np.mha_standard(g["c_q"], g["c_k"], g["c_v"], 8, g["a_wo"], g["a_bo"])
    assert rel_err(o, g["c_out"]) < 1e-5
    o = decoder_np.mha_query_specific(g["b_q"], g["b_k"], g["b_v"], 8, g["a_wo"], g["a_bo"], key_padding_mask=g["b_kpm"])
    assert rel_err(o, g["b_out"]) < 1e-5


@pytest.mark.parametrize("tag", ["a", "b"])
def test_msda_oracle_vs_grid_sample_anchor(tag):
    g = load_golden("msda")
    N, M, D, Lq, P, seed = (int(v) for v in g[f"{tag}_kw"])
    d = synth.make_msda_inputs(N, g[f"{tag}_shapes"], M=M, D=D, Lq=Lq, P=P, seed=seed)
    out = msda_np.msda3d_forward(d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], dt=np.float64)
    assert rel_err(out, g[f"{tag}_out"]) < 1e-5   # fp32 index contract vs fp64 grid_sample coordinates
    out32 = msda_np.msda3d_forward(d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], dt=np.float32)
    assert rel_err(out32, g[f"{tag}_out"]) < 1e-5
    gv, gl, ga = msda_np.msda3d_backward(d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"], g[f"{tag}_go"])
    assert rel_err(gv, g[f"{tag}_gvalue"]) < 1e-5
    assert rel_err(ga, g[f"{tag}_gattn"]) < 1e-5
    assert rel_err(gl, g[f"{tag}_gloc"]) < 1e-4


def test_msda_indices_properties():
    d = synth.make_msda_inputs(1, [(2, 3, 4), (1, 2, 2)], M=2, D=4, Lq=9, P=3, seed=3)
    tl, hl, wl, mask = msda_np.msda3d_indices(d["shapes"], d["loc"])
    assert mask.dtype == np.uint8 and tl.dtype == np.int32
    assert (mask == 0).any() and (mask == 255).any() or (mask != 0).any()
    # a point exactly on a voxel centre reads only corner v1 with weight 1
    loc = np.zeros((1, 1, 1, 1, 1, 3), dtype=np.float32)
    loc[..., 0] = (1 + 0.5) / 4; loc[..., 1] = (2 + 0.5) / 3; loc[..., 2] = (0 + 0.5) / 2
    tl, hl, wl, mask = msda_np.msda3d_indices(np.array([(2, 3, 4)]), loc)
    assert (tl.item(), hl.item(), wl.item()) == (0, 2, 1)


@pytest.mark.parametrize("name", ["grad_tiny", "grad_tiny_masked", "grad_small_masked"])
def test_torch_restatement_gradients_match_reference_autograd(name):
    """oracle/decoder_torch.py (the host baseline of the TRAINING step) against gradients of the unmodified reference."""
    from oracle import decoder_torch
    g = load_golden(name)
    cfg, B, W, inp = case_from_golden(g)
    seed = int(g["meta"][8])
    lw = synth.make_loss_weights(cfg, B, seed=seed)
    loss, grads, gmem, gtgt, gref = decoder_torch.train_step(W, inp, lw, cfg["layers"])
    assert abs(loss - float(g["loss"])) < 1e-4 * max(1.0, abs(float(g["loss"])))
    G = float(np.median([np.abs(g[k]).max() for k in g if k.startswith(("g.", "gs."))]))
    for nm, got in (("memory", gmem), ("tgt", gtgt), ("refpoints_unsigmoid", gref)):
        if "gin." + nm in g:
            assert np.abs(got - g["gin." + nm]).max() <= 1e-4 * max(np.abs(g["gin." + nm]).max(), 1e-2 * G), nm
    for k in g:
        if k.startswith("g."):
            nm = k[2:]
            assert np.abs(grads[nm] - g[k]).max() <= 1e-4 * max(np.abs(g[k]).max(), 1e-2 * G), nm
        elif k.startswith("gs."):
            nm = k[3:]
            idx = synth.grad_sample_index(grads[nm].size, seed)
            assert np.abs(grads[nm].reshape(-1)[idx] - g[k]).max() <= 1e-4 * max(np.abs(g[k]).max(), 1e-2 * G), nm


ENC_CASES = ["enc_tiny", "enc_small_masked", "enc_mid_f2048"]


def enc_case(g):
    B, F_, P, seed, masked = (int(v) for v in g["meta"])
    shapes = [tuple(int(x) for x in r) for r in g["shapes"]]
    W = synth.make_encoder_layer_weights(F_, len(shapes), P, seed=seed)
    inp = synth.make_encoder_inputs(B, shapes, seed=seed, masked=bool(masked))
    return W, inp, shapes, bool(masked)


@pytest.mark.parametrize("name", ENC_CASES)
def test_encoder_layer_oracle_matches_reference(name):
    """oracle/encoder_np.py against the reference DeformableTransformerEncoderLayer + MSDeformAttn3D module + get_reference_points."""
    from oracle import encoder_np
    g = load_golden(name)
    W, inp, shapes, masked = enc_case(g)
    refp = encoder_np.reference_points(shapes, inp["valid_ratios"])
    assert np.abs(refp - g["reference_points"]).max() < 1e-6
    out, attn_out = encoder_np.encoder_layer(W, inp["src"], inp["pos"], refp, shapes, g["level_start"], inp["mask"] if masked else None)
    assert np.abs(attn_out - g["attn_out"]).max() <= 2e-5 * np.abs(g["attn_out"]).max()
    assert np.abs(out - g["out"]).max() <= 2e-5 * np.abs(g["out"]).max()


INTERP_CASES = ["interp_2d_eff", "interp_2d_all", "interp_3d_eff", "interp_3d_all"]


def interp_case(g):
    B, nf, eff, seed = (int(v) for v in g["meta"])
    shapes = [tuple(int(x) for x in r) for r in g["shapes"]]
    rs = np.random.RandomState(6000 + seed)
    Len = sum(t * h * w for t, h, w in shapes)
    tokens = rs.standard_normal((B, Len, 256)).astype(np.float32)
    pos_tokens = rs.standard_normal((B, Len, 256)).astype(np.float32)
    return tokens, pos_tokens, shapes, nf, bool(eff)


@pytest.mark.parametrize("name", INTERP_CASES)
def test_interp_oracle_matches_reference(name):
    """oracle/interp_np.py against the reference's make_interpolated_features + the stack / slice / rearrange of Transformer.forward."""
    from oracle import interp_np
    g = load_golden(name)
    tokens, pos_tokens, shapes, nf, eff = interp_case(g)
    mem = interp_np.interp_to_decoder(tokens, shapes, g["level_start"], nf, eff)
    assert mem.shape == g["memory"].shape
    assert np.abs(mem - g["memory"]).max() <= 2e-5 * np.abs(g["memory"]).max()
    pos0 = interp_np.pos_to_decoder(pos_tokens, shapes, g["level_start"], nf, eff)
    assert np.array_equal(pos0, g["pos0"])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_input_proj_oracle_matches_torch_modules(tag):
    """oracle/encoder_np.input_proj_1x1_gn against nn.Conv3d(k=1) + nn.GroupNorm(32) as models/model.py:64-71 builds them."""
    from oracle import encoder_np
    from oracle.make_golden_inputproj import CASES, make_case
    g = load_golden("inputproj")
    x, w, b, gm, be = make_case(CASES[tag])
    got = encoder_np.input_proj_1x1_gn(x, w, b, gm, be)
    assert np.abs(got - g[tag + "_tokens"]).max() <= 2e-5 * np.abs(g[tag + "_tokens"]).max()


@pytest.mark.parametrize("tag", ["d", "e"])
def test_input_proj_extra_level_oracle_matches_torch_modules(tag):
    """oracle/encoder_np.input_proj_3x3s2_gn against nn.Conv3d(k=3, stride=(1,2,2), padding=1) + nn.GroupNorm(32) (model.py:72-76)."""
    from oracle import encoder_np
    from oracle.make_golden_inputproj import CASES3, make_case3
    g = load_golden("inputproj")
    x, w, b, gm, be = make_case3(CASES3[tag])
    got = encoder_np.input_proj_3x3s2_gn(x, w, b, gm, be)
    assert got.shape == g[tag + "_tokens"].shape
    assert np.abs(got - g[tag + "_tokens"]).max() <= 2e-5 * np.abs(g[tag + "_tokens"]).max()
