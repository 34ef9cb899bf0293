"""Diagnostic (not a test): where does the bf16 path's error come from?  Per-layer errors against the fp32 oracle for
(A) the bf16 engine, (B) the fp32 engine fed bf16-rounded weights (weight-quantisation share)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, decoder_np
from class_query_vad_b200 import DecoderEngine

def rel(a, b): return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-6))

def main(cfgname="ava_vitb", B=1, seed=0, layers=None):
    cfg = dict(synth.CONFIGS[cfgname])
    if layers: cfg["layers"] = layers
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=seed)
    inp = synth.make_decoder_inputs(cfg, B, seed=seed)
    hs, cls_hs, refs = decoder_np.decoder_forward(W, inp["tgt"], inp["memory"], inp["mask"], inp["pos"], inp["refpoints_unsigmoid"], inp["orig_res"], cfg["layers"])
    lg, bx, lb = decoder_np.detr_heads(W, hs, cls_hs, refs)
    t = lambda a: torch.from_numpy(a).cuda()
    def run(Wx, dtype):
        eng = DecoderEngine(Wx, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=dtype, device="cuda")
        o = eng.forward(t(inp["tgt"]), t(inp["memory"]), t(inp["mask"]), t(inp["pos"]), t(inp["refpoints_unsigmoid"]), inp["orig_res"])
        torch.cuda.synchronize()
        return {k: v.float().cpu().numpy() for k, v in o.items()}
    Wq = {k: (torch.from_numpy(v).bfloat16().float().numpy() if (v.ndim >= 2 and v.shape[0] > 8) else v) for k, v in W.items()}
    for tag, o in (("bf16 engine", run(W, torch.bfloat16)), ("fp32 engine, bf16-rounded weights", run(Wq, torch.float32))):
        print(f"== {tag} ({cfgname}, B={B})")
        for l in range(cfg["layers"]):
            print(f"  layer {l}: hs {rel(o['hs'][l], hs[l]):.4f} cls_hs {rel(o['cls_hs'][l], cls_hs[l]):.4f} refs {rel(o['refs'][l], refs[l]):.4f} "
                  f"logits {rel(o['pred_logits'][l], lg[l]):.4f} (|logits|max {np.abs(lg[l]).max():.3f}) boxes {rel(o['pred_boxes'][l], bx[l]):.4f} logits_b {rel(o['pred_logits_b'][l], lb[l]):.4f}", flush=True)

if __name__ == "__main__":
    main("ava_vitb", 1)
    main("small", 3, seed=2)
