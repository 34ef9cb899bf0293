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


if __name__ == "__main__":
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
