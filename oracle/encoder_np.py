"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the deformable encoder layer around the MSDA-3D op.  Never imported by the
product package; tests/, __graft_entry__.smoke() and bench.py's host legs only.  Pinned to the reference by
tests/golden/enc_*.npz (oracle/make_golden_encoder.py; tests/test_oracle_golden.py).

Follows models/detr/dab_transformer.py:433-452 (get_reference_points), :484-523 (DeformableTransformerEncoderLayer) and
ops/modules/ms_deform_attn.py:167-203 (MSDeformAttn3D.forward); the sampling core is oracle/msda_np.py.
"""
import numpy as np

from .msda_np import msda3d_forward


def _ln(x, g, b, eps=1e-5):
    mu = x.mean(-1, keepdims=True)
    var = ((x - mu) ** 2).mean(-1, keepdims=True)
    return (x - mu) / np.sqrt(var + eps) * g + b


def reference_points(shapes, valid_ratios):
    """dab_transformer.py:433-452: voxel centres / (valid_ratio * extent) per level, (x, y, t) order, then scaled by every
    level's valid ratio -> [B, Len, L, 3]."""
    out = []
    for lvl, (T, H, W) in enumerate(np.asarray(shapes).tolist()):
        t, y, x = np.meshgrid(np.linspace(0.5, T - 0.5, T, dtype=np.float32), np.linspace(0.5, H - 0.5, H, dtype=np.float32),
                              np.linspace(0.5, W - 0.5, W, dtype=np.float32), indexing="ij")
        rt = t.reshape(-1)[None] / (valid_ratios[:, None, lvl, 2] * T)
        ry = y.reshape(-1)[None] / (valid_ratios[:, None, lvl, 1] * H)
        rx = x.reshape(-1)[None] / (valid_ratios[:, None, lvl, 0] * W)
        out.append(np.stack((rx, ry, rt), -1))
    ref = np.concatenate(out, 1)
    return (ref[:, :, None] * valid_ratios[:, None]).astype(np.float32)


def msda_module(W, query, refp, src, shapes, level_start, mask=None, M=8, pre="self_attn."):
    """ops/modules/ms_deform_attn.py:167-203.  query, src [B, Len, 256]; refp [B, Len, L, 3] -> [B, Len, 256]."""
    B, Lq, C = query.shape
    L = len(shapes)
    P = W[pre + "attention_weights.bias"].shape[0] // (M * L)
    value = src @ W[pre + "value_proj.weight"].T + W[pre + "value_proj.bias"]                       # :181
    if mask is not None:
        value = np.where(mask[..., None], 0.0, value)                                               # :182-183
    value = value.reshape(B, -1, M, C // M).astype(np.float32)
    offs = (query @ W[pre + "sampling_offsets.weight"].T + W[pre + "sampling_offsets.bias"]).reshape(B, Lq, M, L, P, 3)   # :185
    logit = (query @ W[pre + "attention_weights.weight"].T + W[pre + "attention_weights.bias"]).reshape(B, Lq, M, L * P)  # :186
    e = np.exp(logit - logit.max(-1, keepdims=True))
    attn = (e / e.sum(-1, keepdims=True)).reshape(B, Lq, M, L, P)                                   # :187
    sh = np.asarray(shapes, dtype=np.float32)
    norm = np.stack([sh[:, 0], sh[:, 2], sh[:, 1]], -1)          # (T, W, H) against (x, y, t) offsets: reference quirk, :190
    loc = refp[:, :, None, :, None, :] + offs / norm[None, None, None, :, None, :]                  # :191-192
    out = msda3d_forward(value, np.asarray(shapes, dtype=np.int64), np.asarray(level_start, dtype=np.int64),
                         loc.astype(np.float32), attn.astype(np.float32))                           # :198-199
    out = out.reshape(B, Lq, C)
    return out @ W[pre + "output_proj.weight"].T + W[pre + "output_proj.bias"], loc, attn           # :200


def encoder_layer(W, src, pos, refp, shapes, level_start, mask=None):
    """dab_transformer.py:513-523 in eval mode (dropout = identity)."""
    src2, _, _ = msda_module(W, src + pos, refp, src, shapes, level_start, mask)                    # :515
    x = _ln(src + src2, W["norm1.weight"], W["norm1.bias"])                                         # :516-517
    h = np.maximum(x @ W["linear1.weight"].T + W["linear1.bias"], 0.0)                              # :508
    x = _ln(x + h @ W["linear2.weight"].T + W["linear2.bias"], W["norm2.weight"], W["norm2.bias"])  # :509-510
    return x.astype(np.float32), src2.astype(np.float32)


def input_proj_1x1_gn(x, w, b, gamma, beta, eps=1e-5, groups=32):
    """models/model.py:64-71,162-164: Conv3d(C_in, 256, kernel_size=1) -> GroupNorm(32, 256) on x [B, C_in, T, H, W]; returned
    token-major [B, T*H*W, 256] (= .flatten(2).transpose(1, 2) of the reference's output, dab_transformer.py:317)."""
    B, Cin = x.shape[:2]
    xt = x.reshape(B, Cin, -1).transpose(0, 2, 1).astype(np.float64)              # [B, N, Cin]
    y = xt @ w.reshape(w.shape[0], Cin).T.astype(np.float64) + b                      # [B, N, 256]
    C = y.shape[-1]
    yg = y.reshape(B, -1, groups, C // groups)
    mu = yg.mean(axis=(1, 3), keepdims=True)
    var = yg.var(axis=(1, 3), keepdims=True)
    out = ((yg - mu) / np.sqrt(var + eps)).reshape(B, -1, C) * gamma + beta
    return out.astype(np.float32)


def input_proj_3x3s2_gn(x, w, b, gamma, beta, eps=1e-5, groups=32):
    """models/model.py:72-76,166-170: Conv3d(C_in, 256, kernel_size=3, stride=(1, 2, 2), padding=1) -> GroupNorm(32, 256) on
    x [B, C_in, T, H, W], w [256, C_in, 3, 3, 3]; returned token-major [B, T*Ho*Wo, 256]."""
    B, Cin, T, H, W = x.shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    xp = np.zeros((B, Cin, T + 2, H + 2, W + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1, 1:-1] = x
    y = np.zeros((B, T, Ho, Wo, w.shape[0]), dtype=np.float64)
    for kt in range(3):
        for ky in range(3):
            for kx in range(3):
                patch = xp[:, :, kt:kt + T, ky:ky + 2 * Ho - 1:2, kx:kx + 2 * Wo - 1:2]        # [B, Cin, T, Ho, Wo]
                y += np.einsum("bcthw,oc->bthwo", patch, w[:, :, kt, ky, kx].astype(np.float64))
    y = (y + b).reshape(B, -1, w.shape[0])
    C = y.shape[-1]
    yg = y.reshape(B, -1, groups, C // groups)
    mu = yg.mean(axis=(1, 3), keepdims=True)
    var = yg.var(axis=(1, 3), keepdims=True)
    return (((yg - mu) / np.sqrt(var + eps)).reshape(B, -1, C) * gamma + beta).astype(np.float32)
