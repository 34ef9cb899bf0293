"""MSDA-3D pinned to the REFERENCE KERNEL itself: the reference's CUDA extension
`MultiScaleDeformableAttention` (ops/src/cuda/ms_deform_im2col_cuda_t.cuh:374-439, built for sm_100a by oracle/install_ref.py
from a copy of ops/src whose only change is value.type() -> value.scalar_type(), and shipped under the git-ignored
baseline/_ref/) runs on the same GPU, on the same inputs, beside cqvad_msda3d_forward.

  * forward values: <= 1e-5 (fp32) on the ViT-B/224 pyramid (Len 33 320, M 8, D 32, L 4, P 8) and on the fixture shapes;
  * sampling INDICES bit-exact against the reference kernel.  The reference kernel does not output its indices, so they are
    observed through a probe `value`: channel c of a voxel one-hot-encodes its (t, h, w) coordinates (D = T + H + W channels,
    one level, one point, attention weight 1).  Output channel i of the t-section is then the summed trilinear weight of the
    corners the kernel actually READ at t == i: the non-zero pattern of the output IS the set of (valid) corner indices.
    The patterns of the two kernels must be identical, on locations that sit exactly on voxel centres / borders (where
    loc*dim - 0.5 lands within an ulp of an integer, so that a product-rounded-first, an FMA-contracted and a double-precision
    evaluation pick different voxels), and the low corner decoded from the reference output must equal cqvad_msda3d_indices.
    (This probe is what established the contract: the compiled reference kernel evaluates FFMA(loc, dim, -0.5).)
  * the reference BACKWARD is not a gradient -- it is timed in bench.py, never compared.
"""
import numpy as np
import pytest
import torch

from oracle import synth, ref_import

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_ext():
    m = ref_import.import_reference_msda()
    if m is None:
        pytest.skip("baseline/_ref/MultiScaleDeformableAttention.so not built (python oracle/install_ref.py)")
    return m


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return x if dtype is None else x.to(dtype)


def _ours(value, shapes, lsi, loc, attn):
    from class_query_vad_b200 import MSDeformAttnFunction
    return MSDeformAttnFunction.apply(value, shapes, lsi, loc, attn, 64)


def _ref(ext, value, shapes, lsi, loc, attn):
    step = min(value.shape[0], 64)
    return ext.ms_deform_attn_forward(value, shapes, lsi, loc, attn, step)


def test_msda_forward_matches_reference_kernel_vit_pyramid():
    ext = _ref_ext()
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    d = synth.make_msda_inputs(2, shapes, M=8, D=32, P=8, seed=11, spread=0.4)
    args = (t(d["value"]), t(d["shapes"]), t(d["level_start"]), t(d["loc"]), t(d["attn"]))
    ours, ref = _ours(*args), _ref(ext, *args)
    assert ours.shape == ref.shape == (2, 33320, 256)
    err = (ours - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-5, err
    # bf16 values against the reference kernel on the same (bf16-rounded) values in fp32: one output rounding
    vb = args[0].bfloat16()
    ob = _ours(vb, *args[1:]).float()
    rb = _ref(ext, vb.float(), *args[1:])
    assert ((ob - rb).abs().max() / rb.abs().max()).item() < 6e-3


@pytest.mark.parametrize("N,shapes,M,D,P", [(3, [(2, 5, 7), (1, 3, 4)], 2, 16, 4), (1, [(4, 9, 6)], 4, 40, 3),
                                              (2, [(3, 6, 6), (3, 3, 3), (2, 2, 2)], 8, 32, 8)])
def test_msda_forward_matches_reference_kernel_small_shapes(N, shapes, M, D, P):
    ext = _ref_ext()
    d = synth.make_msda_inputs(N, shapes, M=M, D=D, P=P, seed=N + D, spread=0.6)
    args = (t(d["value"]), t(d["shapes"]), t(d["level_start"]), t(d["loc"]), t(d["attn"]))
    ours, ref = _ours(*args), _ref(ext, *args)
    assert ((ours - ref).abs().max() / ref.abs().max()).item() < 1e-5


def _adversarial_locations(T, H, W, Lq, M, rs):
    """[1, Lq, M, 1, 1, 3] (x, y, t): voxel centres (x_im integral), half-voxel borders, the -1 / dim limits of the in-range
    predicate (cuh:428), nextafter neighbours of all of them, and uniform locations in [-0.1, 1.1]."""
    loc = (-0.1 + 1.2 * rs.uniform(size=(1, Lq, M, 1, 1, 3))).astype(np.float32)
    n = Lq // 2
    dims = np.array([W, H, T], dtype=np.float32)
    k = np.stack([rs.randint(-2, int(d_) + 2, size=(n, M)) for d_ in dims], -1).astype(np.float32)
    half = rs.randint(0, 2, size=(n, M, 3)).astype(np.float32) * 0.5
    base = ((k + half) / dims).astype(np.float32)                     # (k + 0.5)/dim -> x_im = k exactly (or k - 0.5)
    nudge = rs.randint(-1, 2, size=(n, M, 3))
    base = np.where(nudge < 0, np.nextafter(base, np.float32(-10)), np.where(nudge > 0, np.nextafter(base, np.float32(10)), base))
    loc[0, :n, :, 0, 0, :] = base.astype(np.float32)
    return loc


@pytest.mark.parametrize("lvl", [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)])
def test_msda_indices_bit_exact_against_reference_kernel(lvl):
    ext = _ref_ext()
    from class_query_vad_b200 import ms_deform_attn_indices
    T, H, W = lvl
    D, M, Lq = T + H + W, 2, 6000
    D8 = (D + 7) // 8 * 8
    rs = np.random.RandomState(T * 1000 + H)
    # probe value: one-hot coordinates
    tt, hh, ww = np.meshgrid(np.arange(T), np.arange(H), np.arange(W), indexing="ij")
    probe = np.zeros((T * H * W, D8), dtype=np.float32)
    idx = np.arange(T * H * W)
    probe[idx, tt.reshape(-1)] = 1.0
    probe[idx, T + hh.reshape(-1)] = 1.0
    probe[idx, T + H + ww.reshape(-1)] = 1.0
    value = np.ascontiguousarray(np.broadcast_to(probe[None, :, None, :], (1, T * H * W, M, D8)))
    loc = _adversarial_locations(T, H, W, Lq, M, rs)
    attn = np.ones((1, Lq, M, 1, 1), dtype=np.float32)
    shapes = t(np.array([lvl], dtype=np.int64))
    lsi = t(np.zeros(1, dtype=np.int64))
    args = (t(value), shapes, lsi, t(loc), t(attn))
    ours = _ours(*args).reshape(Lq, M, D8).cpu().numpy()
    ref = _ref(ext, *args).reshape(Lq, M, D8).cpu().numpy()
    # (1) identical read pattern: a corner is read by one kernel iff it is read by the other
    np.testing.assert_array_equal(ours != 0, ref != 0)
    assert np.abs(ours - ref).max() < 2e-6
    # (2) the low corner decoded from the REFERENCE output equals cqvad_msda3d_indices wherever it is observable
    tl, hl, wl, mask = (x.cpu().numpy().reshape(Lq, M) for x in ms_deform_attn_indices(shapes, args[3]))
    nz = ref != 0
    checked = 0
    for sec0, n, low in ((0, T, tl), (T, H, hl), (T + H, W, wl)):
        sec = nz[:, :, sec0:sec0 + n]
        cnt = sec.sum(-1)
        first = sec.argmax(-1)
        two = cnt == 2                      # both corners of this axis were read with non-zero weight: low = first
        np.testing.assert_array_equal(first[two], low[two])
        checked += int(two.sum())
        one = (cnt == 1)                    # one corner: low == it (fraction 0 or high corner outside) or low == -1
        ok = (low[one] == first[one]) | ((low[one] == -1) & (first[one] == 0))
        assert ok.all()
    assert checked > int((mask != 0).sum())  # on average more than one axis per sampled point is observable
    # (3) points the reference skipped entirely are exactly the ones with an empty corner mask
    np.testing.assert_array_equal(nz.any(-1), mask != 0)
