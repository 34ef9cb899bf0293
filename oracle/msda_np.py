"""TEST INFRASTRUCTURE -- numpy restatement of the reference 3-D multi-scale deformable attention op.

The reference's own tests do not pin this op (ops/test.py is the stale 2-D Deformable-DETR script, and the shipped kernel has
no CPU implementation: ops/src/cpu/ms_deform_attn_cpu.cpp:26,39 throw).  PINNED HERE in two ways: (1) on the GPU against the
REFERENCE KERNEL ITSELF, compiled from ops/src for sm_100a (oracle/install_ref.py -> baseline/_ref; tests/test_msda_ref_gpu.py:
values <= 1e-5 and the read pattern / low-corner indices bit-exact through a one-hot coordinate probe); (2) on the CPU against
torch's 5-D F.grid_sample (the 3-D analogue of the reference's own ms_deform_attn_core_pytorch,
ops/functions/ms_deform_attn_func.py:48-68) in tests/test_oracle_golden.py and tests/golden/msda.npz.  This restatement
follows the CUDA kernel's arithmetic line by line, including the coordinate rounding the COMPILED kernel performs (below).

  msda3d_indices / msda3d_forward   <- ops/src/cuda/ms_deform_im2col_cuda_t.cuh:33-115 (trilinear), :374-439 (kernel)
  msda3d_backward                   <- the mathematical gradient of that forward (autograd spec).  The reference
                                        backward kernel (cuh:118-243, :441-549) is NOT the gradient of its forward
                                         and is deliberately not restated.
"""
import numpy as np


def _coords(loc, shapes, dt):
    """cuh:412-426.  loc [...,L,P,3] (x,y,t); shapes [L,3] (T,H,W).  The source reads `loc_T * spatial_T - 0.5` (float
    product, double literal); nvcc narrows the subtraction back to fp32 (it is exact either way) and CONTRACTS it with the
    product: the compiled kernel executes `FFMA R, loc, dim, -0.5` (cuobjdump -sass of the reference extension built for
    sm_100a, profiles/r02_sass_summary.txt) -- ONE rounding: fl32(loc*dim - 0.5).  The round-1 reading of the source
    (product rounded first) picked a different voxel on locations that sit within half an ulp of a voxel centre; the
    one-hot index probe against the reference kernel (tests/test_msda_ref_gpu.py) caught it.  In numpy the fused result is
    the float64 evaluation (24-bit x <= 11-bit product and the subtraction are exact in double) rounded once to fp32."""
    if dt != np.float32:            # float64 mode (grid_sample cross-checks): plain arithmetic
        T = shapes[:, 0].astype(dt)[:, None]; H = shapes[:, 1].astype(dt)[:, None]; Wd = shapes[:, 2].astype(dt)[:, None]
        return (loc[..., 2].astype(dt) * T - dt(0.5), loc[..., 1].astype(dt) * H - dt(0.5), loc[..., 0].astype(dt) * Wd - dt(0.5))
    f8 = np.float64
    T = shapes[:, 0].astype(f8)[:, None]
    H = shapes[:, 1].astype(f8)[:, None]
    Wd = shapes[:, 2].astype(f8)[:, None]
    loc32 = loc.astype(np.float32)
    w_im = (loc32[..., 0].astype(f8) * Wd - 0.5).astype(np.float32)
    h_im = (loc32[..., 1].astype(f8) * H - 0.5).astype(np.float32)
    t_im = (loc32[..., 2].astype(f8) * T - 0.5).astype(np.float32)
    return t_im, h_im, w_im


def msda3d_indices(shapes, loc, dt=np.float32):
    """Integer part of the sampling (the 'bit-exact' contract): per (n,q,m,l,p) returns int32 t_low,h_low,w_low and
    an 8-bit mask, bit k set iff corner k (order v1..v8 of cuh:63-109: (t,h,w) low/high with w fastest) is read.
    mask == 0 when the point fails the in-range predicate of cuh:428."""
    shapes = np.asarray(shapes)
    t_im, h_im, w_im = _coords(loc, shapes, dt)
    T = shapes[:, 0][:, None]; H = shapes[:, 1][:, None]; Wd = shapes[:, 2][:, None]
    inside = (t_im > -1) & (h_im > -1) & (w_im > -1) & (t_im < T) & (h_im < H) & (w_im < Wd)      # cuh:428
    tl = np.floor(t_im).astype(np.int32); hl = np.floor(h_im).astype(np.int32); wl = np.floor(w_im).astype(np.int32)
    th, hh, wh = tl + 1, hl + 1, wl + 1
    mask = np.zeros(tl.shape, dtype=np.uint8)
    k = 0
    for tt, tok in ((tl, tl >= 0), (th, th <= T - 1)):
        for hc, hok in ((hl, hl >= 0), (hh, hh <= H - 1)):
            for wc, wok in ((wl, wl >= 0), (wh, wh <= Wd - 1)):
                mask |= ((tok & hok & wok & inside).astype(np.uint8) << k)
                k += 1
    return tl, hl, wl, mask


def msda3d_forward(value, shapes, level_start, loc, attn, dt=np.float32):
    """value [N,Len,M,D]; shapes [L,3] (T,H,W); loc [N,Lq,M,L,P,3]; attn [N,Lq,M,L,P] -> [N,Lq,M*D].
    Accumulation order matches the kernel: levels then points (cuh:401-436), corners v1..v8 (cuh:110-113)."""
    shapes = np.asarray(shapes); level_start = np.asarray(level_start)
    N, Len, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    value = value.astype(dt); attn = attn.astype(dt)
    t_im, h_im, w_im = _coords(loc, shapes, dt)
    tl, hl, wl, mask = msda3d_indices(shapes, loc, dt)
    lt = (t_im - tl).astype(dt); lh = (h_im - hl).astype(dt); lw = (w_im - wl).astype(dt)
    ht, hh_, hw_ = (1 - lt).astype(dt), (1 - lh).astype(dt), (1 - lw).astype(dt)
    out = np.zeros((N, Lq, M, D), dtype=dt)
    n_idx = np.arange(N)[:, None, None]
    m_idx = np.arange(M)[None, None, :]
    for l in range(L):
        T, H, Wd = (int(v) for v in shapes[l])
        v_l = value[:, level_start[l]:level_start[l] + T * H * Wd].reshape(N, T, H, Wd, M, D)
        for p in range(P):
            val = np.zeros((N, Lq, M, D), dtype=dt)
            k = 0
            for dt_, wt in ((0, ht), (1, lt)):
                for dh_, wh_ in ((0, hh_), (1, lh)):
                    for dw_, ww_ in ((0, hw_), (1, lw)):
                        ok = ((mask[:, :, :, l, p] >> k) & 1).astype(bool)
                        ti = np.clip(tl[:, :, :, l, p] + dt_, 0, T - 1)
                        hi = np.clip(hl[:, :, :, l, p] + dh_, 0, H - 1)
                        wi = np.clip(wl[:, :, :, l, p] + dw_, 0, Wd - 1)
                        v = v_l[n_idx, ti, hi, wi, m_idx]                                         # [N,Lq,M,D]
                        wgt = (wt[:, :, :, l, p] * wh_[:, :, :, l, p]).astype(dt) * ww_[:, :, :, l, p]
                        val = val + np.where(ok[..., None], (wgt.astype(dt)[..., None] * v).astype(dt), dt(0))
                        k += 1
            out = out + (val * attn[:, :, :, l, p][..., None]).astype(dt)
    return out.reshape(N, Lq, M * D)


def msda3d_backward(value, shapes, level_start, loc, attn, grad_out, dt=np.float64):
    """Analytic gradient of msda3d_forward w.r.t. value, loc, attn (float64 by default).
    d/dloc_x = W * d/dw_im etc. (w_im = loc_x*W - 0.5); points failing the predicate contribute zero."""
    shapes = np.asarray(shapes); level_start = np.asarray(level_start)
    N, Len, M, D = value.shape
    _, Lq, _, L, P, _ = loc.shape
    value = value.astype(dt); attn = attn.astype(dt)
    go = grad_out.astype(dt).reshape(N, Lq, M, D)
    t_im, h_im, w_im = _coords(loc, shapes, np.float32)   # indices from the fp32 contract
    tl, hl, wl, mask = msda3d_indices(shapes, loc, np.float32)
    t_im, h_im, w_im = t_im.astype(dt), h_im.astype(dt), w_im.astype(dt)
    lt = t_im - tl; lh = h_im - hl; lw = w_im - wl
    g_value = np.zeros_like(value)
    g_loc = np.zeros(loc.shape, dtype=dt)
    g_attn = np.zeros(attn.shape, dtype=dt)
    n_idx = np.broadcast_to(np.arange(N)[:, None, None], (N, Lq, M))
    m_idx = np.broadcast_to(np.arange(M)[None, None, :], (N, Lq, M))
    for l in range(L):
        T, H, Wd = (int(v) for v in shapes[l])
        v_l = value[:, level_start[l]:level_start[l] + T * H * Wd].reshape(N, T, H, Wd, M, D)
        gv_l = np.zeros_like(v_l)
        for p in range(P):
            a = attn[:, :, :, l, p]
            k = 0
            sval = np.zeros((N, Lq, M), dtype=dt)
            for dt_ in (0, 1):
                for dh_ in (0, 1):
                    for dw_ in (0, 1):
                        ok = ((mask[:, :, :, l, p] >> k) & 1).astype(bool)
                        ft = lt[:, :, :, l, p] if dt_ else 1 - lt[:, :, :, l, p]
                        fh = lh[:, :, :, l, p] if dh_ else 1 - lh[:, :, :, l, p]
                        fw = lw[:, :, :, l, p] if dw_ else 1 - lw[:, :, :, l, p]
                        st, sh, sw = (1.0 if dt_ else -1.0), (1.0 if dh_ else -1.0), (1.0 if dw_ else -1.0)
                        ti = np.clip(tl[:, :, :, l, p] + dt_, 0, T - 1)
                        hi = np.clip(hl[:, :, :, l, p] + dh_, 0, H - 1)
                        wi = np.clip(wl[:, :, :, l, p] + dw_, 0, Wd - 1)
                        v = v_l[n_idx, ti, hi, wi, m_idx]                                         # [N,Lq,M,D]
                        dot = (v * go).sum(-1) * ok                                               # [N,Lq,M]
                        sval += ft * fh * fw * dot
                        g_loc[:, :, :, l, p, 0] += a * Wd * sw * ft * fh * dot
                        g_loc[:, :, :, l, p, 1] += a * H * sh * ft * fw * dot
                        g_loc[:, :, :, l, p, 2] += a * T * st * fh * fw * dot
                        contrib = (a * ft * fh * fw * ok)[..., None] * go
                        np.add.at(gv_l, (n_idx, ti, hi, wi, m_idx), contrib)
                        k += 1
            g_attn[:, :, :, l, p] = sval
        g_value[:, level_start[l]:level_start[l] + T * H * Wd] = gv_l.reshape(N, T * H * Wd, M, D)
    return g_value, g_loc, g_attn
