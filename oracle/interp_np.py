"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the encoder -> decoder data-format step of `Transformer.forward`
(models/detr/dab_transformer.py:349-393): un-flatten the encoder output per level, `make_interpolated_features` (:239-294,
F.grid_sample, align_corners=False, zeros padding) onto the (num_frames, H, W) grid of level -2, key-frame slice (`eff`, :378-382)
and the rearrange into the decoder's `L (H W) (B T) C` layout (:391-392).  Pinned by tests/golden/interp_*.npz
(oracle/make_golden_interp.py runs the reference's own make_interpolated_features)."""
import numpy as np


def _unnorm(c, size):
    # grid_sampler_unnormalize, align_corners=False: ((c + 1) * size - 1) / 2   (ATen GridSampler.h)
    return ((c + np.float32(1)) * np.float32(size) - np.float32(1)) / np.float32(2)


def _lin(n):
    return np.linspace(-1, 1, n, dtype=np.float32)


def interp_to_decoder(tokens, shapes, level_start, num_frames, eff):
    """tokens [B, Len, C] (encoder output, levels concatenated, each (t, h, w)-major) -> memory [L, H*W, B*T', C]
    with (T_t, H, W) = shapes[-2] and T' = 1 (eff: frame num_frames // 2) or num_frames."""
    shapes = [tuple(int(v) for v in s) for s in np.asarray(shapes)]
    B, Len, C = tokens.shape
    L = len(shapes)
    Tt, H, W = shapes[L - 2]
    frames = [num_frames // 2] if eff else list(range(num_frames))
    out = np.zeros((L, H * W, B * len(frames), C), dtype=np.float32)
    dh, dw, dt = _lin(H), _lin(W), _lin(num_frames)
    for l, (Tl, Hl, Wl) in enumerate(shapes):
        f = tokens[:, level_start[l]:level_start[l] + Tl * Hl * Wl].reshape(B, Tl, Hl, Wl, C).astype(np.float32)
        for ti, fr in enumerate(frames):
            for i in range(H):
                for j in range(W):
                    if Tt == num_frames:
                        # :256-268 -- 2-D sampling per frame; the grid is stacked (meshy, meshx), i.e. the ROW coordinate dh[i]
                        # is used as grid_sample's x and the column coordinate dw[j] as its y (reference quirk, kept)
                        x, y = _unnorm(dh[i], Wl), _unnorm(dw[j], Hl)
                        corners_t = [(fr, np.float32(1))]
                    else:
                        # :270-283 -- 3-D sampling, grid (x, y, t) = (dw[j], dh[i], dt[fr])
                        x, y, t = _unnorm(dw[j], Wl), _unnorm(dh[i], Hl), _unnorm(dt[fr], Tl)
                        t0 = int(np.floor(t)); lt = np.float32(t - t0)
                        corners_t = [(t0, np.float32(1) - lt), (t0 + 1, lt)]
                    x0, y0 = int(np.floor(x)), int(np.floor(y))
                    lx, ly = np.float32(x - x0), np.float32(y - y0)
                    acc = np.zeros((B, C), dtype=np.float32)
                    for (tz, wt) in corners_t:
                        for (yz, wy) in ((y0, np.float32(1) - ly), (y0 + 1, ly)):
                            for (xz, wx) in ((x0, np.float32(1) - lx), (x0 + 1, lx)):
                                if 0 <= tz < Tl and 0 <= yz < Hl and 0 <= xz < Wl:
                                    acc += (wt * wy * wx) * f[:, tz, yz, xz]
                    out[l, i * W + j, ti::len(frames)] = acc          # (B T) index = b * T' + t'
    return out


def pos_to_decoder(pos_tokens, shapes, level_start, num_frames, eff):
    """lvl_pos_embed_flatten [B, Len, C] -> pos[0] of the decoder [H*W, B*T', C]: level -2, repeated in time (:285,:290), key-frame
    slice, no resampling."""
    shapes = [tuple(int(v) for v in s) for s in np.asarray(shapes)]
    B, Len, C = pos_tokens.shape
    L = len(shapes)
    Tt, H, W = shapes[L - 2]
    frames = [num_frames // 2] if eff else list(range(num_frames))
    p = pos_tokens[:, level_start[L - 2]:level_start[L - 2] + Tt * H * W].reshape(B, Tt, H * W, C)
    out = np.zeros((H * W, B * len(frames), C), dtype=np.float32)
    for ti, fr in enumerate(frames):
        out[:, ti::len(frames)] = p[:, fr % Tt].transpose(1, 0, 2)
    return out
