"""Pipeline trace of one cqvad_mlp launch (CTA 0): python tools/trace_mlp.py M F [relu|gelu] [ln]   (dev tool, MLP_TRACE in mlp_tc.cu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from class_query_vad_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
M, F = int(sys.argv[1]), int(sys.argv[2])
act = 2 if (len(sys.argv) > 3 and sys.argv[3] == "gelu") else 1
ln = len(sys.argv) > 4
X = torch.randn((M, 256), device=dev).bfloat16()
W1 = (torch.randn((F, 256), device=dev) / 16).bfloat16(); b1 = torch.randn(F, device=dev)
W2 = (torch.randn((256, F), device=dev) / F ** 0.5).bfloat16(); b2 = torch.randn(256, device=dev)
g, bt = torch.ones(256, device=dev), torch.zeros(256, device=dev)
Y = torch.empty_like(X); hid = torch.empty((8,), device=dev, dtype=torch.bfloat16)
tr = torch.zeros(4096, dtype=torch.int64, device=dev)
os.environ["CQVAD_MLP_TRACE"] = str(tr.data_ptr())
for it in range(3):
    tr.zero_()
    _lib.check(lib.cqvad_mlp(_lib.BF16, _lib.ptr(X), _lib.ptr(W1), _lib.ptr(b1), _lib.ptr(W2), _lib.ptr(b2), act, _lib.ptr(X),
                             _lib.ptr(g) if ln else None, _lib.ptr(bt) if ln else None, 1e-5, _lib.ptr(Y), _lib.ptr(hid), M, F, _lib.stream_ptr()))
    torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = t[t > 0].min()
def seg(a, n):
    v = t[a:a + n]; return [int(x - t0) if x > 0 else None for x in v]
nch = F // 64
n = min(2 * nch, 128)
print("chunks per tile", nch)
print("producer W1(j) issue          :", seg(0, n))
print("issuer step j begins          :", seg(256, n))
print("issuer hacc_empty(j) passed   :", seg(512, n))
print("issuer w_full for G1(j) passed:", seg(768, n))
print("issuer hs_full for G2(c) passed:", seg(1024, n))
print("issuer w_full for G2(c) passed:", seg(1280, n))
for gi, base in ((0, 1536), (1, 2560)):
    print(f"epi group {gi} chunk wait begins :", seg(base, n // 2))
    print(f"epi group {gi} hacc_full passed  :", seg(base + 256, n // 2))
    print(f"epi group {gi} chunk done        :", seg(base + 512, n // 2))
    for tl in range(3):
        print(f"epi group {gi} tile {tl} Y epilogue [chunks done, y_full passed, LN pass 1 done, stats exchanged, stores issued, residual landed]:", seg(base + 768 + 8 * tl, 6))
