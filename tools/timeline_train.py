"""Per-stream timeline of one decoder training step from the library's profiler scopes (run on the GPU box):
    python tools/timeline_train.py [B]
Prints, per stream, busy time and the largest idle gaps with the classes around them, and how much of the step has 1 / 2 / 3
streams busy at the same time."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from class_query_vad_b200 import DecoderEngine, _lib
from oracle import synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["ava_vitb"]
W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=0)
eng = DecoderEngine(W, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=torch.bfloat16, device=dev, out_f32=False)
inp = synth.make_decoder_inputs("ava_vitb", B, seed=0)
d = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).to(dev) for k in ("tgt", "memory", "pos", "mask", "refpoints_unsigmoid")}
lw = synth.make_loss_weights(cfg, B, seed=1)
g = [torch.from_numpy(lw[k]).to(dev) for k in ("w_hs", "w_cls", "w_refs")]
g[0], g[1] = g[0].bfloat16(), g[1].bfloat16()
lib = _lib.lib()


def step(p=0.1, seed=1):
    eng.forward_train(d["tgt"], d["memory"], d["mask"], d["pos"], d["refpoints_unsigmoid"], (cfg["h"], cfg["w"]), dropout_p=p, seed=seed)
    eng.backward(g[0], g[1], g[2], zero=True, named=False)


for _ in range(3):
    step()
torch.cuda.synchronize()
lib.cqvad_profile_enable(1)
step()
torch.cuda.synchronize()
N = 8192
cls = (ctypes.c_int * N)(); sid = (ctypes.c_int * N)(); t0 = (ctypes.c_double * N)(); t1 = (ctypes.c_double * N)()
n = lib.cqvad_profile_timeline(cls, sid, t0, t1, N)
lib.cqvad_profile_enable(0)
n = min(n, N)
names = [lib.cqvad_profile_class_name(c).decode() for c in range(lib.cqvad_profile_num_classes())]
ev = sorted([(t0[i], t1[i], sid[i], names[cls[i]]) for i in range(n)])
end = max(e[1] for e in ev)
print(f"{n} scopes, span {end:.2f} ms")
for s in sorted(set(e[2] for e in ev)):
    mine = [e for e in ev if e[2] == s]
    busy = sum(e[1] - e[0] for e in mine)
    gaps = sorted(((b[0] - a[1], a, b) for a, b in zip(mine, mine[1:]) if b[0] - a[1] > 0.02), reverse=True)
    print(f"stream {s}: {len(mine)} scopes, busy {busy:.2f} ms, first {mine[0][0]:.2f}, last end {mine[-1][1]:.2f}, idle gaps > 20 us: {len(gaps)} "
          f"totalling {sum(x[0] for x in gaps):.2f} ms")
    for gdt, a, b in gaps[:8]:
        print(f"     gap {gdt * 1e3:7.1f} us at {a[1]:7.3f} ms between [{a[3][:40]}] and [{b[3][:40]}]")
# concurrency histogram
pts = sorted([(e[0], 1) for e in ev] + [(e[1], -1) for e in ev])
lvl, last, hist = 0, 0.0, {}
for t, dlt in pts:
    hist[lvl] = hist.get(lvl, 0.0) + (t - last)
    lvl += dlt; last = t
print("time with k scopes open (scopes of one stream can nest at most 1 deep):", {k: round(v, 2) for k, v in sorted(hist.items())})
per = {}
for a, b, s, nm in ev:
    per[nm] = per.get(nm, 0.0) + (b - a)
for nm, v in sorted(per.items(), key=lambda kv: -kv[1]):
    print(f"  {v:7.3f} ms  {nm}")
if os.environ.get("HEAD"):
    k = int(os.environ["HEAD"])
    print("first scopes:")
    for a, b, s, nm in ev[:k]:
        print(f"   s{s} {a:7.3f} +{(b - a) * 1e3:7.1f} us  {nm}")
    print("last scopes:")
    for a, b, s, nm in ev[-k:]:
        print(f"   s{s} {a:7.3f} +{(b - a) * 1e3:7.1f} us  {nm}")
