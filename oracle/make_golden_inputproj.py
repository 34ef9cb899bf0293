"""TEST INFRASTRUCTURE ONLY.  tests/golden/inputproj.npz: `input_proj[l]` of the reference DETR head for the non-ViT backbones
(models/model.py:64-71: nn.Sequential(nn.Conv3d(in_channels, hidden_dim, kernel_size=1), nn.GroupNorm(32, hidden_dim)), applied at
:162-164) followed by the flatten of Transformer.forward (dab_transformer.py:317).  The two modules are stock torch.nn layers; they
are instantiated exactly as the reference does (importing models/model.py itself needs yacs / timm / the video backbones) with
seeded non-degenerate parameters.  Run anywhere torch is available:   python -m oracle.make_golden_inputproj"""
import os
import numpy as np
import torch
from torch import nn

from .make_golden import GOLD

CASES = {"a": dict(B=2, Cin=512, shape=(2, 5, 6), seed=0), "b": dict(B=1, Cin=2048, shape=(4, 7, 7), seed=1),
         "c": dict(B=3, Cin=64, shape=(1, 3, 3), seed=2)}


def make_case(kw):
    rs = np.random.RandomState(12000 + kw["seed"])
    x = rs.standard_normal((kw["B"], kw["Cin"]) + kw["shape"]).astype(np.float32)
    w = (rs.standard_normal((256, kw["Cin"], 1, 1, 1)) / np.sqrt(kw["Cin"])).astype(np.float32)
    b = (0.3 * rs.standard_normal(256)).astype(np.float32)
    g = (1.0 + 0.2 * rs.standard_normal(256)).astype(np.float32)
    be = (0.1 * rs.standard_normal(256)).astype(np.float32)
    return x, w, b, g, be


CASES3 = {"d": dict(B=2, Cin=64, shape=(2, 7, 7), seed=3), "e": dict(B=1, Cin=128, shape=(3, 6, 5), seed=4)}


def make_case3(kw):
    rs = np.random.RandomState(13000 + kw["seed"])
    x = rs.standard_normal((kw["B"], kw["Cin"]) + kw["shape"]).astype(np.float32)
    w = (rs.standard_normal((256, kw["Cin"], 3, 3, 3)) / np.sqrt(27 * kw["Cin"])).astype(np.float32)
    b = (0.3 * rs.standard_normal(256)).astype(np.float32)
    g = (1.0 + 0.2 * rs.standard_normal(256)).astype(np.float32)
    be = (0.1 * rs.standard_normal(256)).astype(np.float32)
    return x, w, b, g, be


def main():
    out = {}
    for tag, kw in CASES3.items():                     # the extra level, models/model.py:72-76
        x, w, b, g, be = make_case3(kw)
        proj = nn.Sequential(nn.Conv3d(kw["Cin"], 256, kernel_size=3, stride=(1, 2, 2), padding=1), nn.GroupNorm(32, 256))
        with torch.no_grad():
            proj[0].weight.copy_(torch.from_numpy(w)); proj[0].bias.copy_(torch.from_numpy(b))
            proj[1].weight.copy_(torch.from_numpy(g)); proj[1].bias.copy_(torch.from_numpy(be))
            y5 = proj(torch.from_numpy(x))
        out[tag + "_tokens"] = y5.flatten(2).transpose(1, 2).contiguous().numpy()
        out[tag + "_shape"] = np.array(y5.shape[2:], dtype=np.int64)
    for tag, kw in CASES.items():
        x, w, b, g, be = make_case(kw)
        proj = nn.Sequential(nn.Conv3d(kw["Cin"], 256, kernel_size=1), nn.GroupNorm(32, 256))      # models/model.py:67-70
        with torch.no_grad():
            proj[0].weight.copy_(torch.from_numpy(w)); proj[0].bias.copy_(torch.from_numpy(b))
            proj[1].weight.copy_(torch.from_numpy(g)); proj[1].bias.copy_(torch.from_numpy(be))
            y = proj(torch.from_numpy(x)).flatten(2).transpose(1, 2).contiguous()                     # dab_transformer.py:317
        out[tag + "_tokens"] = y.numpy()
    np.savez_compressed(os.path.join(GOLD, "inputproj.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
