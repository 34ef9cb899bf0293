"""Forward bf16 parity figures per output tensor (run on the GPU box): relative L2 and max-norm against the reference fixtures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from helpers import load_golden, case_from_meta
from test_decoder_gpu import run_engine, CASES

for name in CASES:
    g = load_golden(name)
    cfg, B, W, inp = case_from_meta(g["meta"])
    out, _ = run_engine(cfg, W, inp, torch.bfloat16)
    row = []
    for k in ("hs", "cls_hs", "refs", "pred_boxes", "pred_logits_b", "pred_logits"):
        ref = g.get(k)
        got = out[k]
        if ref is None and k == "cls_hs":
            ref = g["cls_hs_sub"]; got = got[:, :, ::4, ::7, ::5]
        l2 = np.linalg.norm(got.astype(np.float64) - ref) / np.linalg.norm(ref)
        mx = np.abs(got - ref).max() / np.abs(ref).max()
        row.append(f"{k} L2 {l2:.2e} max {mx:.2e}")
    print(name, " | ".join(row), flush=True)
