"""Times cqvad_linear (tcgen05 GEMM + epilogue) on the decoder's large shapes: python tools/bench_gemm.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from class_query_vad_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
SHAPES = [(94080, 1024, 256, 0), (94080, 512, 256, 0), (94080, 256, 256, 0), (94080, 256, 1024, 0), (38400, 2048, 256, 1),
          (38400, 256, 2048, 0), (94080, 256, 512, 0), (94080, 256, 256, 3)]
for (M, N, K, mode) in SHAPES:
    A = torch.randn((M, K), device=dev).bfloat16()
    W = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev)
    C = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    res = torch.randn((M, N), device=dev).bfloat16() if mode == 3 else None
    act = 1 if mode == 1 else 0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = lib.cqvad_linear(_lib.BF16, _lib.ptr(A), _lib.ptr(W), _lib.ptr(b), _lib.ptr(res), _lib.ptr(C), M, N, K, act, _lib.stream_ptr())
        e1.record()
        torch.cuda.synchronize()
        _lib.check(rc)
        ts.append(e0.elapsed_time(e1) * 1e3)
    t = sorted(ts)[len(ts) // 2]
    byt = (M * K + M * N + N * K) * 2 + (M * N * 2 if res is not None else 0)
    ref = (A.float() @ W.float().t() + b)
    if act: ref = ref.relu()
    if res is not None: ref = ref + res.float()
    err = float((C.float() - ref).abs().max() / ref.abs().max())
    print(f"M={M} N={N} K={K} mode={mode}: {t:7.1f} us  {2.0 * M * N * K / t / 1e6:7.1f} TFLOP/s  {byt / t / 1e3:6.0f} GB/s(alg)  relerr {err:.1e}", flush=True)
