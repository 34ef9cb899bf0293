"""Times one deformable encoder layer (cqvad_deform_encoder_layer_forward) on the ViT-B/224 pyramid: python tools/bench_encoder.py [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from class_query_vad_b200 import pack_encoder_layer_weights, encoder_layer_forward
from oracle import synth, encoder_np

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
shapes = [(8, 56, 56), (8, 28, 28), (8, 14, 14), (8, 7, 7)]
F_, P = 2048, 8
dev = torch.device("cuda:0")
W = synth.make_encoder_layer_weights(F_, 4, P, seed=5)
inp = synth.make_encoder_inputs(1, shapes, seed=5)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
packed = pack_encoder_layer_weights({k: torch.from_numpy(v) for k, v in W.items()}, torch.bfloat16, dev)
refp = t(encoder_np.reference_points(shapes, inp["valid_ratios"])).repeat(B, 1, 1, 1)
sh = torch.tensor(shapes, dtype=torch.int64, device=dev)
ls = torch.cat((sh.new_zeros((1,)), sh.prod(1).cumsum(0)[:-1]))
src = t(inp["src"]).bfloat16().repeat(B, 1, 1).contiguous()
pos = t(inp["pos"]).bfloat16().repeat(B, 1, 1).contiguous()
Len = src.shape[1]
iters = int(os.environ.get("ITERS", "6"))
ts = []
for it in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = encoder_layer_forward(packed, src, pos, refp, sh, ls, None, P, F_)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
rows = B * Len
gflop = 2.0 * rows * 256 * (256 + 768 + 256 + 256 + 2 * F_) / 1e9          # linears only (the sampling is a gather)
print(f"B={B} Len={Len}: {ms:.3f} ms per layer  ({B / ms * 1e3:.1f} clips/s/layer, {gflop / ms:.1f} TFLOP/s of GEMM work, {gflop / B:.1f} GFLOP/clip)")

# ---- training step of the layer (forward keeping activations + backward) through the autograd Function ----
from class_query_vad_b200 import DeformableTransformerEncoderLayer
layer = DeformableTransformerEncoderLayer(d_model=256, d_ffn=F_, n_levels=4, n_heads=8, n_points=P)
layer.load_state_dict({k: torch.from_numpy(v) for k, v in W.items()}, strict=True)
layer = layer.to(dev).eval()
srcg = src.clone().requires_grad_(True)
posg = pos.clone().requires_grad_(True)
go = torch.randn_like(src)
tf, tb = [], []
for it in range(iters):
    for p_ in layer.parameters():
        p_.grad = None
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    out = layer(srcg, posg, refp, sh, ls, None)
    e1.record()
    out.backward(go)
    e2.record()
    torch.cuda.synchronize()
    tf.append(e0.elapsed_time(e1)); tb.append(e1.elapsed_time(e2))
mf, mb = sorted(tf)[len(tf) // 2], sorted(tb)[len(tb) // 2]
print(f"B={B} training step of one layer: fwd {mf:.3f} ms + bwd {mb:.3f} ms = {mf + mb:.3f} ms  ({B / (mf + mb) * 1e3:.1f} clips/s/layer, "
      f"{3 * gflop / (mf + mb):.1f} TFLOP/s of GEMM work)")
