"""Achieved HBM GB/s of the memory-bound kernels (CUDA events, warm-up, inputs > L2 or L2 flushed between iterations).
Prints one JSON line per kernel: algorithmic bytes (DESIGN.md section 4 / BASELINE.md section 3) / time vs MEASURED_PEAKS.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from class_query_vad_b200 import MSDeformAttnFunction
from class_query_vad_b200.modules.ops import layer_norm
from oracle import synth

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()                      # 256 MB write: evicts the 126 MB L2
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, nbytes, ms, note=""):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "achieved_GBs": round(gbs, 1),
                      "peak_GBs": PEAK, "frac": round(gbs / PEAK, 3), "note": note}), flush=True)


def main():
    B = 4
    shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]           # ViT-B/224 pyramid, Len = 33 320
    d = synth.make_msda_inputs(B, shapes, M=8, D=32, P=8, seed=1, spread=0.15)
    t = lambda a: torch.from_numpy(a).cuda()
    sh, ls, loc, attn = t(d["shapes"]), t(d["level_start"]), t(d["loc"]), t(d["attn"])
    for dt, es in ((torch.float32, 4), (torch.bfloat16, 2)):
        value = t(d["value"]).to(dt)
        Len = value.shape[1]
        # value + loc (fp32) + attn (fp32) + out
        nbytes = B * Len * 256 * es * 2 + B * Len * 8 * 32 * 4 * 4
        ms = timeit(lambda: MSDeformAttnFunction.apply(value, sh, ls, loc, attn, 64))
        report(f"msda3d_forward[{str(dt).split('.')[-1]}] B={B} Len={Len}", nbytes, ms, "loc/attn fp32")
    value = t(d["value"]).requires_grad_(True); locg = loc.clone().requires_grad_(True); attg = attn.clone().requires_grad_(True)
    out = MSDeformAttnFunction.apply(value, sh, ls, locg, attg, 64)
    go = torch.randn_like(out)
    Len = value.shape[1]
    nb_bwd = B * Len * 256 * 4 * (1 + 1 + 2) + B * Len * 8 * 32 * 4 * 4 * 2   # value, grad_out, grad_value RMW; loc+attn read, grads written
    ms = timeit(lambda: torch.autograd.grad(out, (value, locg, attg), go, retain_graph=True))
    report(f"msda3d_backward[float32] B={B} Len={Len}", nb_bwd, ms, "includes zero-fill of grad_value by the caller")
    rows = 94080
    x = torch.randn(rows, 256, device="cuda").bfloat16(); g = torch.ones(256, device="cuda"); b = torch.zeros(256, device="cuda")
    ms = timeit(lambda: layer_norm(x, g, b))
    report(f"layernorm[bf16] rows={rows}", rows * 256 * 2 * 2, ms)
    xf = torch.randn(rows, 256, device="cuda")
    ms = timeit(lambda: layer_norm(xf, g, b))
    report(f"layernorm[float32] rows={rows}", rows * 256 * 4 * 2, ms)


if __name__ == "__main__":
    main()
