"""TEST INFRASTRUCTURE -- numpy restatement of the reference position encodings.  Not imported by the product.

  * position_embedding_sine_3d   <- models/position_encoding.py:32-73 (PositionEmbeddingSine_3D.forward, normalize=True)
  * gen_sineembed_for_position   <- models/detr/dab_transformer.py:50-76
"""
import math
import numpy as np


def position_embedding_sine_3d(mask, num_pos_feats=256, temperature=10000.0, dtype=np.float32):
    """mask [B,T,H,W] bool (True = padded) -> [B, num_pos_feats, T, H, W].

    Channel split (position_encoding.py:22-23): t gets num_pos_feats/8*2 (=64), y and x num_pos_feats/8*3 (=96).
    The exponent uses TRUE division `torch.div(i, 2)` (position_encoding.py:55,60), not floor division, so
    consecutive sin/cos channels have different frequencies."""
    f = dtype
    not_mask = ~mask
    t_embed = np.cumsum(not_mask, axis=1, dtype=np.float32).astype(f)
    y_embed = np.cumsum(not_mask, axis=2, dtype=np.float32).astype(f)
    x_embed = np.cumsum(not_mask, axis=3, dtype=np.float32).astype(f)
    eps = f(1e-6)
    scale = f(2 * math.pi)
    t_embed = t_embed / (t_embed[:, -1:, :, :] + eps) * scale          # :47
    y_embed = y_embed / (y_embed[:, :, -1:, :] + eps) * scale          # :48
    x_embed = x_embed / (x_embed[:, :, :, -1:] + eps) * scale          # :49
    n_t = num_pos_feats / 8 * 2
    n_s = num_pos_feats / 8 * 3
    t_dim = np.arange(int(n_t), dtype=f)
    t_dim = np.power(f(temperature), (2 * (t_dim / 2) / f(n_t)).astype(f)).astype(f)   # :55
    s_dim = np.arange(int(n_s), dtype=f)
    s_dim = np.power(f(temperature), (2 * (s_dim / 2) / f(n_s)).astype(f)).astype(f)   # :60
    pos_t = t_embed[..., None] / t_dim
    pos_x = x_embed[..., None] / s_dim
    pos_y = y_embed[..., None] / s_dim

    def interleave(p):                                                  # :65-69
        return np.stack((np.sin(p[..., 0::2]), np.cos(p[..., 1::2])), axis=5).reshape(p.shape)

    pos = np.concatenate((interleave(pos_t), interleave(pos_y), interleave(pos_x)), axis=4)   # :71
    return np.ascontiguousarray(pos.transpose(0, 4, 1, 2, 3)).astype(f)


def gen_sineembed_for_position(pos_tensor, dtype=np.float32):
    """[nq, BT, 4] (x, y, w, h) in [0,1] -> [nq, BT, 512] ordered (y, x, w, h), 128 dims each
    (dab_transformer.py:50-76); dim_t = 10000 ** (2*(i//2)/128) (floor division here)."""
    f = dtype
    pos_tensor = pos_tensor.astype(f)
    scale = f(2 * math.pi)
    i = np.arange(128, dtype=f)
    dim_t = np.power(f(10000.0), (2 * np.floor(i / 2) / f(128)).astype(f)).astype(f)

    def emb(col):
        p = (pos_tensor[:, :, col] * scale)[:, :, None] / dim_t
        return np.stack((np.sin(p[:, :, 0::2]), np.cos(p[:, :, 1::2])), axis=3).reshape(p.shape)

    return np.concatenate((emb(1), emb(0), emb(2), emb(3)), axis=2).astype(f)
