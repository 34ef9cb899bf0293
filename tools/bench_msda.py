"""MSDA-3D kernels timed alone on the ViT-B/224 pyramid (Len = Lq = 33 320, M 8, D 32, L 4, P 8), B clips, against the HBM
roofline (algorithmic bytes of BASELINE.md section 3) and beside the REFERENCE kernel (baseline/_ref, same inputs).

Sampling locations are the ones a deformable ENCODER produces (ops/modules/ms_deform_attn.py:187-192): reference point of the
query's own voxel + offsets / (T_l, W_l, H_l) with offsets ~ N(0, sigma voxels) -- i.e. spatially coherent between neighbouring
queries; `--uniform` draws them uniformly over the volume instead (no locality at all, the worst case for any cache).
CUDA events, L2 flushed (256 MB write) before every timed launch, median of 10.

  python tools/bench_msda.py [--B 4] [--sigma 1.75] [--uniform] [--old]      (--old: round-1 forward kernel, CQVAD_MSDA_OLD_FWD)
"""
import argparse
import json
import os
import sys

ap = argparse.ArgumentParser()
ap.add_argument("--B", type=int, default=4)
ap.add_argument("--sigma", type=float, default=1.75)
ap.add_argument("--uniform", action="store_true")
ap.add_argument("--old", action="store_true")
ap.add_argument("--no-ref", action="store_true")
args = ap.parse_args()
if args.old:
    os.environ["CQVAD_MSDA_OLD_FWD"] = "1"
    os.environ["CQVAD_MSDA_OLD_BWD"] = "1"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from class_query_vad_b200 import _lib  # noqa: E402
from oracle import ref_import  # noqa: E402

pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def encoder_like_inputs(B, shapes, M, D, P, sigma, uniform, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    L = len(shapes)
    Len = sum(t * h * w for t, h, w in shapes)
    refs = []
    for (T, H, W) in shapes:      # get_reference_points (dab_transformer.py:433-452), valid ratios 1
        t, y, x = torch.meshgrid(torch.arange(T, device=dev) + 0.5, torch.arange(H, device=dev) + 0.5, torch.arange(W, device=dev) + 0.5,
                                 indexing="ij")
        refs.append(torch.stack((x.reshape(-1) / W, y.reshape(-1) / H, t.reshape(-1) / T), -1))
    ref = torch.cat(refs, 0)                                                      # [Len, 3] (x, y, t)
    if uniform:
        loc = torch.rand((B, Len, M, L, P, 3), device=dev, generator=g)
    else:
        off = sigma * torch.randn((B, Len, M, L, P, 3), device=dev, generator=g)
        norm = torch.tensor([[T, W, H] for (T, H, W) in shapes], dtype=torch.float32, device=dev)   # the reference's (T, W, H) order
        loc = ref[None, :, None, None, None, :] + off / norm[None, None, None, :, None, :]
    attn = torch.softmax(torch.randn((B, Len, M, L * P), device=dev, generator=g), -1).reshape(B, Len, M, L, P)
    value = torch.randn((B, Len, M, D), device=dev, generator=g)
    sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
    ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
    return value, sh, ls, loc.contiguous(), attn.contiguous(), Len


def main():
    B, M, D, P = args.B, 8, 32, 8
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
    value, sh, ls, loc, attn, Len = encoder_like_inputs(B, shapes, M, D, P, args.sigma, args.uniform)
    L = len(shapes)
    lib = _lib.lib()
    p = _lib.ptr
    st = _lib.stream_ptr
    tag = ("uniform locations" if args.uniform else f"encoder-like locations (sigma {args.sigma} voxels)") + (", round-1 kernels" if args.old else "")
    ref_ext = None if args.no_ref else ref_import.import_reference_msda()
    res = []
    for dt, es in ((torch.bfloat16, 2), (torch.float32, 4)):
        v = value.to(dt).contiguous()
        out = torch.empty((B, Len, M * D), dtype=dt, device=dev)
        nbytes = B * Len * M * D * es * 2 + B * Len * M * L * P * 4 * 4       # value + out, loc (3) + attn (1) fp32
        ms = timeit(lambda: _lib.check(lib.cqvad_msda3d_forward(_lib.dtype_id(dt), p(v), p(sh), p(ls), p(loc), p(attn), p(out),
                                                                B, Len, M, D, L, Len, P, st())))
        res.append(dict(kernel=f"cqvad_msda3d_forward[{str(dt).split('.')[-1]}]", ms=round(ms, 4), algorithmic_MB=round(nbytes / 1e6, 1),
                        achieved_GBs=round(nbytes / ms / 1e6, 1), frac_hbm=round(nbytes / ms / 1e6 / PEAK, 3)))
        go = torch.randn((B, Len, M * D), device=dev).to(dt)
        gv = torch.zeros((B, Len, M, D), dtype=torch.float32, device=dev)
        gl = torch.empty_like(loc)
        ga = torch.empty_like(attn)
        nb_b = B * Len * M * D * (es * 2 + 4 * 2) + B * Len * M * L * P * 4 * 4 * 2   # value, grad_out read; grad_value fp32 RMW; loc/attn read + grads written

        def bwd():
            gv.zero_()
            _lib.check(lib.cqvad_msda3d_backward(_lib.dtype_id(dt), p(v), p(sh), p(ls), p(loc), p(attn), p(go), p(gv), p(gl), p(ga),
                                                 B, Len, M, D, L, Len, P, st()))
        ms = timeit(bwd)
        res.append(dict(kernel=f"cqvad_msda3d_backward[{str(dt).split('.')[-1]}] (+ grad_value zero-fill)", ms=round(ms, 4),
                        algorithmic_MB=round(nb_b / 1e6, 1), achieved_GBs=round(nb_b / ms / 1e6, 1), frac_hbm=round(nb_b / ms / 1e6 / PEAK, 3)))
    if ref_ext is not None:
        v = value.contiguous()
        nbytes = B * Len * M * D * 4 * 2 + B * Len * M * L * P * 4 * 4
        ms = timeit(lambda: ref_ext.ms_deform_attn_forward(v, sh, ls, loc, attn, min(B, 64)))
        res.append(dict(kernel="REFERENCE ms_deform_attn_forward[float32] (baseline/_ref)", ms=round(ms, 4), algorithmic_MB=round(nbytes / 1e6, 1),
                        achieved_GBs=round(nbytes / ms / 1e6, 1), frac_hbm=round(nbytes / ms / 1e6 / PEAK, 3)))
        go = torch.randn((B, Len, M * D), device=dev)
        try:
            ms = timeit(lambda: ref_ext.ms_deform_attn_backward(v, sh, ls, loc, attn, go, min(B, 64)))
            res.append(dict(kernel="REFERENCE ms_deform_attn_backward[float32] (baseline/_ref; not a gradient, timed only)", ms=round(ms, 4)))
        except Exception as e:
            res.append(dict(kernel="REFERENCE ms_deform_attn_backward", error=str(e)[:120]))
    for r in res:
        r["case"] = f"B={B} Len={Len} {tag}"
        r["peak_GBs"] = PEAK
        print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
