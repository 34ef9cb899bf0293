"""Prints per-tensor gradient errors of the CUDA training path against the reference-autograd fixtures (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import load_golden, case_from_golden
from test_train_gpu import run_train, grad_errors

names = sys.argv[2:] or ["grad_tiny"]
dtype = torch.float32 if sys.argv[1] == "f32" else torch.bfloat16
for name in names:
    g = load_golden(name)
    cfg, B, W, inp = case_from_golden(g)
    seed = int(g["meta"][8])
    loss, grads, eng = run_train(cfg, B, W, inp, seed, dtype)
    print(name, "loss", loss, "ref", float(g["loss"]), "launches fwd", eng.last_launches, "bwd", eng.last_launches_bwd)
    errs, l2 = grad_errors(grads, g, seed, tgt_zero=bool(int(g["meta"][10])), want_l2=True)
    print("   rel-L2: median", float(np.median(list(l2.values()))), "max", max(l2.items(), key=lambda kv: kv[1]), "p90", float(np.percentile(list(l2.values()), 90)))
    for k, v in sorted(errs.items(), key=lambda kv: -kv[1] if np.isfinite(kv[1]) else -1e30)[:int(os.environ.get("TOPN", "40"))]:
        print(f"   {v:10.3e}  {k}")
    print("   median", float(np.median(list(errs.values()))), "n", len(errs))
