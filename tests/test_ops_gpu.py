"""GPU parity of the individual C-ABI ops against the CPU oracle / golden fixtures / plain torch fp32 references."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_err, TOL_FP32, TOL_BF16
from oracle import synth, msda_np, posenc_np, decoder_np

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


# ---------------------------------------------------------------- MSDA ------------------------------------------------
def _msda_case(tag):
    g = load_golden("msda")
    N, M, D, Lq, P, seed = (int(v) for v in g[f"{tag}_kw"])
    d = synth.make_msda_inputs(N, g[f"{tag}_shapes"], M=M, D=D, Lq=Lq, P=P, seed=seed)
    return g, d


@pytest.mark.parametrize("tag", ["a", "b"])
def test_msda_forward_backward_vs_golden(tag):
    from class_query_vad_b200 import MSDeformAttnFunction
    g, d = _msda_case(tag)
    value = t(d["value"]).requires_grad_(True)
    loc = t(d["loc"]).requires_grad_(True)
    attn = t(d["attn"]).requires_grad_(True)
    out = MSDeformAttnFunction.apply(value, t(d["shapes"]), t(d["level_start"]), loc, attn, 64)
    assert rel_err(out.detach().cpu().numpy(), g[f"{tag}_out"]) < 1e-5
    (out * t(g[f"{tag}_go"], torch.float32)).sum().backward()
    assert rel_err(value.grad.cpu().numpy(), g[f"{tag}_gvalue"]) < 1e-4
    assert rel_err(attn.grad.cpu().numpy(), g[f"{tag}_gattn"]) < 1e-4
    assert rel_err(loc.grad.cpu().numpy(), g[f"{tag}_gloc"]) < 1e-3


def test_msda_indices_bit_exact_and_vit_shape():
    """Bit-exact integer contract on the ViT-B/224 pyramid (Len 33 320) with locations that hit every border case."""
    from class_query_vad_b200 import ms_deform_attn_indices, MSDeformAttnFunction
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    d = synth.make_msda_inputs(1, shapes, M=8, D=32, Lq=3000, P=8, seed=4, spread=0.4)
    loc = d["loc"]
    # adversarial locations: exact voxel centres / borders, where loc*dim - 0.5 sits on an integer
    T, H, W = 8, 56, 56
    loc[0, :50, 0, 0, :, 0] = (np.arange(50)[:, None] % W + 0.5) / W
    loc[0, :50, 0, 0, :, 1] = (np.arange(50)[:, None] % H + 0.5) / H
    loc[0, :50, 0, 0, :, 2] = (np.arange(8)[None, :] + 0.5) / T
    loc[0, 50:60, :, :, :, :] = 0.0
    loc[0, 60:70, :, :, :, :] = 1.0
    tl, hl, wl, mask = ms_deform_attn_indices(t(d["shapes"]), t(loc))
    otl, ohl, owl, omask = msda_np.msda3d_indices(d["shapes"], loc)
    np.testing.assert_array_equal(tl.cpu().numpy(), otl)
    np.testing.assert_array_equal(hl.cpu().numpy(), ohl)
    np.testing.assert_array_equal(wl.cpu().numpy(), owl)
    np.testing.assert_array_equal(mask.cpu().numpy(), omask)
    out = MSDeformAttnFunction.apply(t(d["value"]), t(d["shapes"]), t(d["level_start"]), t(loc), t(d["attn"]), 64)
    ref = msda_np.msda3d_forward(d["value"], d["shapes"], d["level_start"], loc, d["attn"])
    assert rel_err(out.cpu().numpy(), ref) < 1e-5
    # bf16 values, fp32 locations
    outb = MSDeformAttnFunction.apply(t(d["value"], torch.bfloat16), t(d["shapes"]), t(d["level_start"]), t(loc), t(d["attn"]), 64)
    assert rel_err(outb.float().cpu().numpy(), ref) < TOL_BF16


def test_msda_linearity_full_size():
    """Size-independent property at the BASELINE shape (Lq = Len = 33 320): the op is linear in `value` and in `attn`."""
    from class_query_vad_b200 import MSDeformAttnFunction
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    d = synth.make_msda_inputs(1, shapes, M=8, D=32, P=8, seed=5)
    f = lambda v, a: MSDeformAttnFunction.apply(v, t(d["shapes"]), t(d["level_start"]), t(d["loc"]), a, 64)
    v1, a1 = t(d["value"]), t(d["attn"])
    v2 = torch.roll(v1, 1, dims=1)
    o1, o2, o12 = f(v1, a1), f(v2, a1), f(v1 + 2 * v2, a1)
    assert rel_err((o1 + 2 * o2).cpu().numpy(), o12.cpu().numpy()) < 1e-5
    assert rel_err(f(v1, 3 * a1).cpu().numpy(), (3 * o1).cpu().numpy()) < 1e-5
    assert o1.shape == (1, 33320, 256)


def test_msda_edge_cases():
    from class_query_vad_b200 import MSDeformAttnFunction
    shapes = torch.tensor([[1, 2, 2]], dtype=torch.int64, device=DEV)
    lsi = torch.zeros(1, dtype=torch.int64, device=DEV)
    # empty query set
    out = MSDeformAttnFunction.apply(torch.zeros(1, 4, 1, 4, device=DEV), shapes, lsi,
                                     torch.zeros(1, 0, 1, 1, 1, 3, device=DEV), torch.zeros(1, 0, 1, 1, 1, device=DEV), 64)
    assert out.shape == (1, 0, 4)
    # all points out of range -> exact zeros; batch not a multiple of im2col_step is fine
    v = torch.randn(3, 4, 1, 4, device=DEV)
    loc = torch.full((3, 5, 1, 1, 2, 3), 7.0, device=DEV)
    out = MSDeformAttnFunction.apply(v, shapes, lsi, loc, torch.ones(3, 5, 1, 1, 2, device=DEV), 2)
    assert out.abs().max().item() == 0.0
    with pytest.raises(RuntimeError, match="contiguous"):
        MSDeformAttnFunction.apply(v.transpose(0, 1).contiguous().transpose(0, 1), shapes, lsi, loc, torch.ones(3, 5, 1, 1, 2, device=DEV), 2)


# ---------------------------------------------------------------- GEMM ------------------------------------------------
GEMM_SHAPES = [(128, 256, 64), (128, 256, 256), (480, 256, 256), (1000, 512, 256), (300, 1024, 256), (257, 256, 1024),
               (130, 2048, 256), (513, 256, 2048), (96, 128, 128), (64, 8, 64), (480, 256, 512)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_linear_bf16_tensor_core(M, N, K, act):
    from class_query_vad_b200.modules.ops import linear
    rs = np.random.RandomState(M + N + K)
    x = t(rs.standard_normal((M, K)).astype(np.float32), torch.bfloat16)
    w = t((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32), torch.bfloat16)
    b = t(rs.standard_normal(N).astype(np.float32))
    r = t(rs.standard_normal((M, N)).astype(np.float32), torch.bfloat16)
    got = linear(x, w, b, act=act, res=r).float()
    ref = x.float() @ w.float().T + b
    ref = torch.relu(ref) if act == 1 else (torch.nn.functional.gelu(ref) if act == 2 else ref)
    ref = ref + r.float()
    assert rel_err(got.cpu().numpy(), ref.cpu().numpy()) < 1e-2   # one bf16 rounding of the output


@pytest.mark.parametrize("M,N,K", [(100, 256, 256), (77, 24, 128), (200, 1024, 256)])
def test_linear_fp32(M, N, K):
    from class_query_vad_b200.modules.ops import linear
    rs = np.random.RandomState(1)
    x = t(rs.standard_normal((M, K)).astype(np.float32)); w = t((rs.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32))
    b = t(rs.standard_normal(N).astype(np.float32))
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = torch.nn.functional.gelu(x.double() @ w.double().T + b.double())
    assert rel_err(linear(x, w, b, act=2).cpu().numpy(), ref.cpu().numpy()) < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, TOL_BF16)])
@pytest.mark.parametrize("n,h,w", [(3, 14, 14), (2, 16, 16), (5, 3, 5), (1, 1, 1), (2, 7, 128)])
def test_convblock(dtype, tol, n, h, w):
    """ConvBlock (3x3 conv as implicit GEMM on the y-padded layout + LN + MLP + residual) vs the numpy oracle."""
    from class_query_vad_b200.modules.ops import conv_block
    W = synth.make_decoder_weights(5, 1, 128, seed=3)
    p = "cls_layers.0.conv_blocks.0."
    rs = np.random.RandomState(n * 100 + h)
    x = rs.standard_normal((n, h, w, 256)).astype(np.float32)
    ref = decoder_np.conv_block(W, p, x)
    got = conv_block(t(x, dtype), *(t(W[p + k]) for k in ("conv1.weight", "conv1.bias", "norm.weight", "norm.bias",
                                                         "conv2.weight", "conv2.bias", "conv3.weight", "conv3.bias")))
    assert rel_err(got.float().cpu().numpy(), ref) < tol


# ---------------------------------------------------------------- attention / norms / pos-enc -----------------------
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, TOL_BF16)])
def test_multihead_attention_modes_vs_reference_golden(dtype, tol):
    from class_query_vad_b200 import MultiheadAttention
    g = load_golden("attention")

    def mk(E, vdim, qsk=False):
        m = MultiheadAttention(E, 8, dropout=0.1, vdim=vdim, query_specific_key=qsk).eval()
        m.out_proj.weight.data = torch.from_numpy(g["a_wo"]); m.out_proj.bias.data = torch.from_numpy(g["a_bo"])
        return m.to(DEV)

    o, w_ = mk(256, 256)(t(g["a_q"], dtype), t(g["a_k"], dtype), t(g["a_v"], dtype), key_padding_mask=t(g["a_kpm"]))
    assert w_ is None and rel_err(o.float().cpu().numpy(), g["a_out"]) < tol
    o, _ = mk(512, 256)(t(g["c_q"], dtype), t(g["c_k"], dtype), t(g["c_v"], dtype))
    assert rel_err(o.float().cpu().numpy(), g["c_out"]) < tol
    o, _ = mk(512, 256, True)(t(g["b_q"], dtype), t(g["b_k"], dtype), t(g["b_v"], dtype), key_padding_mask=t(g["b_kpm"]))
    assert rel_err(o.float().cpu().numpy(), g["b_out"]) < tol


def test_layernorm_and_posenc_vs_golden():
    from class_query_vad_b200 import PositionEmbeddingSine_3D, gen_sineembed_for_position
    from class_query_vad_b200.modules.ops import layer_norm
    g = load_golden("posenc")

    class NT:
        def __init__(self, m): self.tensors, self.mask = None, m
    pe = PositionEmbeddingSine_3D(256, normalize=True)
    assert rel_err(pe(NT(t(g["mask"]))).cpu().numpy(), g["pos"]) < 1e-4
    pos2 = pe(NT(torch.zeros(1, 8, 14, 14, dtype=torch.bool, device=DEV))).cpu().numpy()
    assert rel_err(pos2[:, :, ::3, ::5, ::4], g["pos_vit14"]) < 1e-4
    assert rel_err(gen_sineembed_for_position(t(g["ref_in"])).cpu().numpy(), g["sine"]) < 1e-4
    rs = np.random.RandomState(0)
    x = rs.standard_normal((37, 256)).astype(np.float32) * 3 + 1
    gam = rs.standard_normal(256).astype(np.float32); bet = rs.standard_normal(256).astype(np.float32)
    ref = decoder_np.layer_norm(x, gam, bet, 1e-6)
    assert rel_err(layer_norm(t(x), t(gam), t(bet), 1e-6).cpu().numpy(), ref) < 1e-5


def test_module_level_layers_fp32():
    """Per-layer drop-ins (TransformerDecoderLayer / TransformerClassDecoderLayer forward signatures) vs oracle taps."""
    from class_query_vad_b200 import build_decoder, gen_sineembed_for_position
    g = load_golden("dec_tiny")
    from helpers import case_from_meta
    cfg, B, W, inp = case_from_meta(g["meta"])
    taps = {}
    decoder_np.decoder_forward(W, inp["tgt"], inp["memory"], inp["mask"], inp["pos"], inp["refpoints_unsigmoid"],
                               inp["orig_res"], cfg["layers"], taps=taps)
    dec = build_decoder(cfg["nq"], cfg["K"], cfg["layers"], cfg["F"])
    dec.load_state_dict({k: torch.from_numpy(v) for k, v in W.items() if not k.startswith("heads.")}, strict=True)
    dec = dec.to(DEV).eval()
    ref = torch.sigmoid(t(inp["refpoints_unsigmoid"]))
    sine = gen_sineembed_for_position(ref)
    query_pos = dec.ref_point_head(sine)
    qse = t(taps["l0.qse"])
    out, actor, qmem = dec.layers[0](t(inp["tgt"]), t(inp["memory"]), memory_key_padding_mask=t(inp["mask"]), pos=t(inp["pos"]),
                                     query_pos=query_pos, query_sine_embed=qse, is_first=True)
    assert rel_err(out.cpu().numpy(), g["l0.output"]) < TOL_FP32
    assert rel_err(actor.cpu().numpy(), g["l0.actor"]) < TOL_FP32
    assert rel_err(qmem.cpu().numpy(), taps["l0.q_memory"]) < TOL_FP32
    cls_out, nxt = dec.cls_layers[0](actor, qmem, t(inp["pos"])[0], qse, dec.class_queries.weight, inp["orig_res"], cfg["nq"], True)
    assert rel_err(cls_out.cpu().numpy(), g["l0.cls_output"]) < TOL_FP32


# ---- weight-gradient GEMM (tcgen05, MN-major operands) -----------------------------------------------------------------
def _wgrad(dY, X, N, K, conv=None, use_ws=True):
    from class_query_vad_b200 import _lib
    lib = _lib.lib()
    dev = dY.device
    dW = torch.zeros((N, 9 * K) if conv else (N, K), dtype=torch.float32, device=dev)
    db = torch.zeros(N, dtype=torch.float32, device=dev)
    ws = torch.empty(lib.cqvad_wgrad_workspace_bytes(), dtype=torch.uint8, device=dev) if use_ws else None
    h, w = conv if conv else (0, 0)
    rc = lib.cqvad_linear_wgrad(_lib.dtype_id(dY.dtype), _lib.ptr(dY), _lib.ptr(X), _lib.ptr(dW), _lib.ptr(db), dY.shape[0], N, K, h, w,
                                _lib.ptr(ws), 0 if ws is None else ws.numel(), _lib.stream_ptr())
    _lib.check(rc)
    torch.cuda.synchronize()
    return dW, db


@pytest.mark.parametrize("M,N,K", [(4096, 256, 256), (5000, 1024, 256), (3011 * 8, 256, 1024), (2048, 128, 256), (2304, 512, 256),
                                   (7777 * 8, 256, 2048)])
def test_wgrad_tc_matches_fp32(M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N)
    dY = torch.randn((M, N), device="cuda", generator=g).bfloat16()
    X = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    dW, db = _wgrad(dY, X, N, K)
    ref = dY.double().t() @ X.double()
    assert float((dW.double() - ref).abs().max() / ref.abs().max()) < 1e-4     # bf16-exact inputs, fp32 accumulation
    assert float((db.double() - dY.double().sum(0)).abs().max() / dY.double().sum(0).abs().max()) < 1e-4
    dW2, db2 = _wgrad(dY, X, N, K, use_ws=False)                                  # CUDA-core kernel, same contract
    assert float((dW2.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    # accumulate semantics: a second call adds
    from class_query_vad_b200 import _lib
    lib = _lib.lib()
    ws = torch.empty(lib.cqvad_wgrad_workspace_bytes(), dtype=torch.uint8, device="cuda")
    rc = lib.cqvad_linear_wgrad(_lib.BF16, _lib.ptr(dY), _lib.ptr(X), _lib.ptr(dW), _lib.ptr(db), M, N, K, 0, 0, _lib.ptr(ws), ws.numel(),
                                _lib.stream_ptr())
    _lib.check(rc)
    assert float((dW.double() - 2 * ref).abs().max() / ref.abs().max()) < 2e-4


@pytest.mark.parametrize("n_img,h,w,dt", [(40, 14, 14, torch.bfloat16), (9, 16, 16, torch.bfloat16), (70, 3, 5, torch.bfloat16),
                                          (6, 6, 7, torch.float32), (150, 7, 9, torch.bfloat16)])
def test_conv_wgrad_matches_torch(n_img, h, w, dt):
    """dW of the 3x3 conv on the y-padded NHWC layout against torch's conv2d weight gradient (fp64)."""
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(n_img)
    C = 256
    x = torch.randn((n_img, h, w, C), device="cuda", generator=g).to(dt)
    dy = torch.randn((n_img, h, w, C), device="cuda", generator=g).to(dt)
    pad = lambda t: torch.cat([t, torch.zeros((n_img, 1, w, C), device="cuda", dtype=dt)], 1).reshape(-1, C).contiguous()
    dW, db = _wgrad(pad(dy), pad(x), C, C, conv=(h, w))
    xr = x.double().permute(0, 3, 1, 2).requires_grad_(False)
    wt = torch.zeros((C, C, 3, 3), dtype=torch.float64, device="cuda", requires_grad=True)
    out = F.conv2d(xr, wt, padding=1)
    out.backward(dy.double().permute(0, 3, 1, 2))
    ref = wt.grad.permute(0, 2, 3, 1).reshape(C, 9 * C)            # [O][ky*3+kx][I]
    assert float((dW.double() - ref).abs().max() / ref.abs().max()) < 1e-4
    assert float((db.double() - dy.double().sum((0, 1, 2))).abs().max()) / float(dy.double().sum((0, 1, 2)).abs().max()) < 1e-4


@pytest.mark.parametrize("M,N,K,dt", [(20000, 1024, 256, torch.bfloat16), (300, 512, 256, torch.bfloat16), (19001, 1024, 256, torch.bfloat16),
                                      (777, 96, 64, torch.float32)])
def test_linear_gelu_train_dual_epilogue(M, N, K, dt):
    """act = gelu(A W^T + b), dact = gelu'(A W^T + b) from ONE epilogue (CTA-pair / single-CTA tcgen05 and CUDA-core paths)
    against torch (nn.GELU erf form, dab_transformer.py:84) in fp64."""
    from class_query_vad_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = torch.randn((M, K), device="cuda", generator=g).to(dt)
    W = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5 * 1.5).to(dt)
    b = torch.randn(N, device="cuda", generator=g)
    act = torch.empty((M, N), device="cuda", dtype=dt)
    dact = torch.empty_like(act)
    _lib.check(lib.cqvad_linear_gelu_train(_lib.dtype_id(dt), _lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(act), _lib.ptr(dact),
                                           M, N, K, _lib.stream_ptr()))
    x = (A.double() @ W.double().t() + b.double()).requires_grad_(True)
    y = torch.nn.functional.gelu(x)
    y.sum().backward()
    tol = 6e-3 if dt == torch.bfloat16 else 2e-5     # one bf16 rounding of a value <= ~5 (2^-8 relative) / fp32 accumulation
    assert float((act.double() - y.detach()).abs().max() / y.detach().abs().max()) < tol
    assert float((dact.double() - x.grad).abs().max()) < tol * 1.2           # |gelu'| <= 1.13


@pytest.mark.parametrize("M,N,K,mode", [(20000, 1024, 256, 3), (20000, 2048, 256, 1), (300, 256, 256, 3), (19001, 512, 256, 1)])
def test_linear_dgrad_act_epilogue(M, N, K, mode):
    """dX = (dY Wt^T) * aux (stored derivative) / masked by aux > 0 (ReLU) in the TMA-store epilogue (side input by TMA)."""
    from class_query_vad_b200 import _lib
    lib = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + mode)
    dY = torch.randn((M, K), device="cuda", generator=g).bfloat16()
    Wt = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).bfloat16()
    aux = torch.randn((M, N), device="cuda", generator=g).bfloat16()
    dX = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    _lib.check(lib.cqvad_linear_dgrad_act(_lib.BF16, _lib.ptr(dY), _lib.ptr(Wt), _lib.ptr(aux), mode, _lib.ptr(dX), M, N, K,
                                          _lib.stream_ptr()))
    ref = dY.double() @ Wt.double().t()
    ref = ref * aux.double() if mode == 3 else ref * (aux.double() > 0)
    assert float((dX.double() - ref).abs().max() / ref.abs().max()) < 6e-3
