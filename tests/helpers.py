"""Shared test helpers: golden loading, error metric of SURVEY.md section 7 (rel = |a-b|_inf / max(|b|_inf, 1e-6))."""
import os
import numpy as np
from oracle import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# tolerances stated by BASELINE.json north_star
TOL_FP32 = 1e-3
TOL_BF16 = 2e-2


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return {k: z[k] for k in z.files}


def case_from_golden(g):
    """case_from_meta + the fixture's stored bias overrides (`wb.<name>`: ReLU-kink clearing, oracle/make_golden_grads.py)."""
    cfg, B, W, inp = case_from_meta(g["meta"])
    for k in g:
        if k.startswith("wb."):
            assert W[k[3:]].shape == g[k].shape, k
            W[k[3:]] = np.ascontiguousarray(g[k], dtype=np.float32)
    return cfg, B, W, inp


def case_from_meta(meta):
    B, nq, tp, h, w, K, layers, F, seed, masked, tgt_zero = (int(v) for v in meta)
    cfg = dict(nq=nq, tprime=tp, h=h, w=w, K=K, layers=layers, F=F)
    W = synth.make_decoder_weights(K, layers, F, seed=seed)
    inp = synth.make_decoder_inputs(cfg, B, seed=seed, masked=bool(masked), tgt_zero=bool(tgt_zero))
    return cfg, B, W, inp
