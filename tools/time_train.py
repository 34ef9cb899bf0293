"""Times forward_train / backward of the decoder at a given batch (run on the GPU box): python tools/time_train.py [B] [layers]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from class_query_vad_b200 import DecoderEngine
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
cfg = dict(synth.CONFIGS["ava_vitb"])
if len(sys.argv) > 2:
    cfg["layers"] = int(sys.argv[2])
dev = torch.device("cuda:0")
W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=0)
eng = DecoderEngine(W, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=torch.bfloat16, device=dev, out_f32=False)
inp = synth.make_decoder_inputs(cfg, B, seed=0)
t = lambda a: torch.from_numpy(a).to(dev)
d = {k: t(inp[k]) for k in ("tgt", "memory", "pos", "mask", "refpoints_unsigmoid")}
lw = synth.make_loss_weights(cfg, B, seed=0)
gh, gc, gr = t(lw["w_hs"]).bfloat16(), t(lw["w_cls"]).bfloat16(), t(lw["w_refs"])
ev = lambda: torch.cuda.Event(enable_timing=True)
for it in range(int(os.environ.get('ITERS', '4'))):
    e0, e1, e2 = ev(), ev(), ev()
    import time
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    out = eng.forward_train(d["tgt"], d["memory"], d["mask"], d["pos"], d["refpoints_unsigmoid"], (cfg["h"], cfg["w"]))
    t1 = time.perf_counter()
    e1.record()
    g = eng.backward(gh, gc, gr, named=False)
    t2 = time.perf_counter()
    e2.record()
    torch.cuda.synchronize()
    print(f"   host enqueue: fwd {1e3 * (t1 - t0):.2f} ms, bwd {1e3 * (t2 - t1):.2f} ms")
    print(f"iter {it}: fwd {e0.elapsed_time(e1):.2f} ms ({eng.last_launches} launches)  bwd {e1.elapsed_time(e2):.2f} ms ({eng.last_launches_bwd} launches)"
          f"  -> {B / (e0.elapsed_time(e2) * 1e-3):.1f} clips/s", flush=True)
import ctypes
if os.environ.get('NOPROF'):
    sys.exit(0)
from class_query_vad_b200 import _lib
lib = _lib.lib()
lib.cqvad_profile_enable(1)
out = eng.forward_train(d["tgt"], d["memory"], d["mask"], d["pos"], d["refpoints_unsigmoid"], (cfg["h"], cfg["w"]))
g = eng.backward(gh, gc, gr, named=False)
torch.cuda.synchronize()
for c in range(lib.cqvad_profile_num_classes()):
    tot, sc, ln = ctypes.c_double(), ctypes.c_long(), ctypes.c_long()
    lib.cqvad_profile_read(c, ctypes.byref(tot), ctypes.byref(sc), ctypes.byref(ln))
    if sc.value:
        print(f"   {tot.value:8.3f} ms  {sc.value:5d} scopes {ln.value:5d} launches  {lib.cqvad_profile_class_name(c).decode()}")
lib.cqvad_profile_enable(0)
print("workspace GB", eng._tws.numel() / 1e9)
