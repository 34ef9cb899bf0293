"""Diagnostic: bf16 decoder on one small case, with CUDA_LAUNCH_BLOCKING semantics (sync after the call) and error print."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import synth, decoder_np
from class_query_vad_b200 import DecoderEngine
for cfgname, B in (("tiny", 2), ("small", 3), ("ava_vitb", 2)):
    cfg = dict(synth.CONFIGS[cfgname]); cfg["layers"] = min(cfg["layers"], 2)
    W = synth.make_decoder_weights(cfg["K"], cfg["layers"], cfg["F"], seed=0)
    inp = synth.make_decoder_inputs(cfg, B, seed=0)
    hs, cls_hs, refs = decoder_np.decoder_forward(W, inp["tgt"], inp["memory"], inp["mask"], inp["pos"], inp["refpoints_unsigmoid"], inp["orig_res"], cfg["layers"])
    t = lambda a: torch.from_numpy(a).cuda()
    eng = DecoderEngine(W, nq=cfg["nq"], K=cfg["K"], layers=cfg["layers"], F=cfg["F"], dtype=torch.bfloat16, device="cuda")
    o = eng.forward(t(inp["tgt"]), t(inp["memory"]), t(inp["mask"]), t(inp["pos"]), t(inp["refpoints_unsigmoid"]), inp["orig_res"])
    torch.cuda.synchronize()
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    print(cfgname, "hs", rel(o["hs"].float().cpu().numpy(), hs), "cls_hs", rel(o["cls_hs"].float().cpu().numpy(), cls_hs), flush=True)
