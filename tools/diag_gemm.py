"""Diagnostic (not a test): error statistics of the tcgen05 GEMM / conv kernels per shape, printed without asserting,
so that one GPU call localises a descriptor / pipeline bug (which rows, which columns, which k-blocks)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from class_query_vad_b200.modules.ops import linear
from class_query_vad_b200 import _lib


def one(M, N, K, act=0, res=False, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    r = torch.randn(M, N, device="cuda", generator=g).bfloat16() if res else None
    ref = x.float() @ w.float().T + b
    if act == 1: ref = torch.relu(ref)
    if act == 2: ref = torch.nn.functional.gelu(ref)
    if res: ref = ref + r.float()
    got = linear(x, w, b, act=act, res=r).float()
    torch.cuda.synchronize()
    err = (got - ref).abs()
    rel = err.max().item() / ref.abs().max().item()
    bad = err > 0.05 * ref.abs().max()
    msg = f"M={M:6d} N={N:5d} K={K:5d} act={act} res={int(res)} rel={rel:.3e} bad={bad.float().mean().item():.4f}"
    if bad.any():
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        msg += f" bad_rows[{rows.numel()}]={rows[:8].tolist()}..{rows[-3:].tolist()} bad_cols[{cols.numel()}]={cols[:8].tolist()}..{cols[-3:].tolist()}"
        msg += f" got[0,:4]={got[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}"
    print(msg, flush=True)
    return rel


if __name__ == "__main__" and len(sys.argv) == 1:
    torch.backends.cuda.matmul.allow_tf32 = False
    print("device", torch.cuda.get_device_name(0))
    for (M, N, K) in [(128, 256, 64), (128, 256, 128), (128, 256, 256), (256, 256, 256), (128, 512, 256), (480, 256, 512),
                      (1000, 1024, 256), (513, 256, 2048), (94080, 256, 256), (38400, 2048, 256)]:
        one(M, N, K)
    one(300, 256, 256, act=1, res=True)
    one(300, 1024, 256, act=2)
    _lib.lib().cqvad_debug_force_simt(1)
    one(300, 256, 256, act=1, res=True)
    _lib.lib().cqvad_debug_force_simt(0)


def one_mlp(M, F, act, res, ln, seed=0):
    from class_query_vad_b200.modules.ops import ffn
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(M, 256, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(F, 256, device="cuda", generator=g) / 16).bfloat16()
    w2 = (torch.randn(256, F, device="cuda", generator=g) / F ** 0.5).bfloat16()
    b1 = torch.randn(F, device="cuda", generator=g) * 0.1; b2 = torch.randn(256, device="cuda", generator=g) * 0.1
    r = x if res else None
    gam = torch.randn(256, device="cuda", generator=g) if ln else None
    bet = torch.randn(256, device="cuda", generator=g) if ln else None
    h = x.float() @ w1.float().T + b1
    h = torch.relu(h) if act == 1 else torch.nn.functional.gelu(h)
    h = h.bfloat16().float()
    ref = h @ w2.float().T + b2
    if res: ref = ref + x.float()
    if ln: ref = torch.nn.functional.layer_norm(ref, (256,), gam, bet, 1e-5)
    got = ffn(x, w1, b1, w2, b2, act=act, res=r, ln_weight=gam, ln_bias=bet).float()
    torch.cuda.synchronize()
    err = (got - ref).abs(); rel = err.max().item() / ref.abs().max().item()
    bad = err > 0.05 * ref.abs().max()
    msg = f"MLP M={M:6d} F={F:5d} act={act} res={int(res)} ln={int(ln)} rel={rel:.3e} bad={bad.float().mean().item():.4f}"
    if bad.any():
        rows = bad.any(1).nonzero().flatten(); cols = bad.any(0).nonzero().flatten()
        msg += f" bad_rows[{rows.numel()}]={rows[:8].tolist()} bad_cols[{cols.numel()}]={cols[:8].tolist()} got[0,:4]={got[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}"
    print(msg, flush=True)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "mlp":
    for (M, F, act, res, ln) in [(128, 256, 1, False, False), (128, 1024, 2, True, False), (300, 2048, 1, True, True),
                                 (1000, 1024, 2, True, False), (38400, 2048, 1, True, True), (100800, 1024, 2, True, False)]:
        one_mlp(M, F, act, res, ln)
