"""Pipeline trace of one cqvad_linear launch (CTA 0): python tools/trace_gemm.py M N K  (dev tool, see CQ_TRACE in gemm_tc.cu)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from class_query_vad_b200 import _lib
lib = _lib.lib()
dev = torch.device("cuda:0")
M, N, K = (int(a) for a in sys.argv[1:4])
A = torch.randn((M, K), device=dev).bfloat16()
W = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
b = torch.randn(N, device=dev)
C = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
tr = torch.zeros(2048, dtype=torch.int64, device=dev)
for it in range(3):
    flush.zero_()
    tr.zero_()
    os.environ["CQVAD_GEMM_TRACE"] = str(tr.data_ptr())
    lib.cqvad_linear(_lib.BF16, _lib.ptr(A), _lib.ptr(W), _lib.ptr(b), None, _lib.ptr(C), M, N, K, 0, _lib.stream_ptr())
    torch.cuda.synchronize()
t = tr.cpu().numpy()
t0 = t[t > 0].min()
def seg(a, b):
    v = t[a:b]; v = v[v > 0]
    return [(int(x - t0)) for x in v]
kb = K // 64
print("producer stage issue (ns):", seg(0, 256)[:6 * kb])
print("mma full-wait done   (ns):", seg(256, 512)[:6 * kb])
print("mma tempty-wait done (ns):", seg(512, 640)[:8])
print("epi tfull-wait done  (ns):", seg(640, 768)[:8])
print("epi tile done        (ns):", seg(768, 896)[:8])
print("last events:", seg(0, 256)[-1], seg(256, 512)[-1], seg(768, 896)[-1], "n tiles", len(seg(768, 896)))
for e in range(1, 4):
    v = t[896 + e * 16: 896 + e * 16 + 13]
    print("epi tile", e, "[barsync, h0:waitrd, ld0, -, ld1, -, pre-fence, stored | h1: waitrd, ld0, -, ld1, -, pre-fence, stored]:", [int(x - v[0]) if x > 0 else None for x in v])
