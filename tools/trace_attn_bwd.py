"""Phase trace of attn_bwd_tc_kernel (CTA (0,0)) inside one training step: python tools/trace_attn_bwd.py  (dev tool, AB_TRACE)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
dev = torch.device("cuda:0")
tr = torch.zeros(64, dtype=torch.int64, device=dev)
os.environ["CQVAD_ATTN_TRACE"] = str(tr.data_ptr())
os.environ["NOPROF"] = "1"; os.environ["ITERS"] = "2"
sys.argv = [sys.argv[0], "32", "1"]
try:
    exec(open(os.path.join(os.path.dirname(__file__), "time_train.py")).read())
except SystemExit:
    pass
torch.cuda.synchronize()
t = tr.cpu().numpy()
names = ["entry", "setup done", "loads landed", "-", "S,dP ready", "pass1 max", "pass2 sum", "pass3 P/dS stored", "dV,dQ,dK ready", "epilogue done", "exit"]
for tag, o in (("cross-attention (4 heads, hd 64)", 0), ("self-attention (8 heads, hd 32, fused)", 16)):
    v = t[o:o + 11]
    print(tag, {n: int(x - v[0]) for n, x in zip(names, v) if x > 0})
