"""TEST INFRASTRUCTURE ONLY.  tests/golden/neck.npz: the ViT simple-feature-pyramid neck of the reference
(models/backbone_3d_builder.py:133-182 `lateral_convs`, applied by `space_forward` :190-200) on synthetic ViT features.
The reference `Backbone` cannot be constructed here (it builds the ViT body and needs timm), so the four `nn.Sequential`s are
built exactly as lines 139-180 build them, with the reference's OWN channel-first `LayerNorm` class (:20-40) executed from the
reference source file at generation time (nothing is copied into this repository).  fp32, CPU.
Run in the build container only:   python -m oracle.make_golden_neck"""
import ast
import os
import numpy as np
import torch
from torch import nn

from .ref_import import REF_ROOT
from .make_golden import GOLD

# fixture cases: (tag, B, Cin, T, H, W)   -- Cin % 256 == 0 keeps every GEMM K a multiple of 64
CASES = {"a": dict(B=1, Cin=256, T=2, H=4, W=4, seed=0), "b": dict(B=1, Cin=768, T=3, H=6, W=5, seed=1)}
SCALES = (4.0, 2.0, 1.0, 0.5)


def reference_layernorm_class():
    src = open(os.path.join(REF_ROOT, "models", "backbone_3d_builder.py")).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LayerNorm")
    ns = {"nn": nn, "torch": torch}
    exec(compile(ast.Module(body=[node], type_ignores=[]), "backbone_3d_builder.py", "exec"), ns)
    return ns["LayerNorm"]


def build_lateral(LayerNorm, dim, scale, out_channel=256):
    """models/backbone_3d_builder.py:139-180 for one scale."""
    if scale == 4.0:
        layers = [nn.ConvTranspose3d(dim, dim // 2, kernel_size=[1, 2, 2], stride=[1, 2, 2]), LayerNorm(dim // 2), nn.GELU(),
                  nn.ConvTranspose3d(dim // 2, dim // 4, kernel_size=[1, 2, 2], stride=[1, 2, 2])]
        out_dim = dim // 4
    elif scale == 2.0:
        layers = [nn.ConvTranspose3d(dim, dim // 2, kernel_size=[1, 2, 2], stride=[1, 2, 2])]
        out_dim = dim // 2
    elif scale == 1.0:
        layers, out_dim = [], dim
    else:
        layers, out_dim = [nn.MaxPool3d(kernel_size=[1, 2, 2], stride=[1, 2, 2])], dim
    layers.extend([nn.Conv3d(out_dim, out_channel, kernel_size=1, bias=False), LayerNorm(out_channel),
                   nn.Conv3d(out_channel, out_channel, kernel_size=3, padding=1, bias=False)])
    return nn.Sequential(*layers)


def make_case(kw):
    """Deterministic input + weights: x [B,Cin,T,H,W]; state dicts of the four levels under the reference's names."""
    rs = np.random.RandomState(9000 + kw["seed"])
    x = rs.standard_normal((kw["B"], kw["Cin"], kw["T"], kw["H"], kw["W"])).astype(np.float32)
    LN = reference_layernorm_class() if os.path.isdir(REF_ROOT) else None
    sds = []
    for scale in SCALES:
        sd = {}
        shapes = {}
        dim = kw["Cin"]
        idx = 0
        if scale == 4.0:
            shapes = {"0.weight": (dim, dim // 2, 1, 2, 2), "0.bias": (dim // 2,), "1.weight": (dim // 2,), "1.bias": (dim // 2,),
                      "3.weight": (dim // 2, dim // 4, 1, 2, 2), "3.bias": (dim // 4,)}
            idx, out_dim = 4, dim // 4
        elif scale == 2.0:
            shapes = {"0.weight": (dim, dim // 2, 1, 2, 2), "0.bias": (dim // 2,)}
            idx, out_dim = 1, dim // 2
        elif scale == 1.0:
            idx, out_dim = 0, dim
        else:
            idx, out_dim = 1, dim
        shapes.update({f"{idx}.weight": (256, out_dim, 1, 1, 1), f"{idx + 1}.weight": (256,), f"{idx + 1}.bias": (256,),
                       f"{idx + 2}.weight": (256, 256, 3, 3, 3)})
        for name, shp in shapes.items():
            if len(shp) == 1:
                v = (1.0 + 0.2 * rs.standard_normal(shp)) if name.endswith("weight") else 0.2 * rs.standard_normal(shp)
            else:
                fan = shp[0] if name.startswith(("0.", "3.")) and len(shp) == 5 and shp[2:] == (1, 2, 2) else int(np.prod(shp[1:]))
                v = rs.standard_normal(shp) / np.sqrt(fan)
            sd[name] = v.astype(np.float32)
        sds.append(sd)
    return x, sds


def main():
    LN = reference_layernorm_class()
    out = {}
    for tag, kw in CASES.items():
        x, sds = make_case(kw)
        for lvl, (scale, sd) in enumerate(zip(SCALES, sds)):
            seq = build_lateral(LN, kw["Cin"], scale)
            seq.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
            with torch.no_grad():
                y = seq(torch.from_numpy(x))
            out[f"{tag}.{lvl}"] = y.numpy()
            print(tag, lvl, tuple(y.shape), float(y.abs().max()))
    np.savez_compressed(os.path.join(GOLD, "neck.npz"), **out)


if __name__ == "__main__":
    main()
