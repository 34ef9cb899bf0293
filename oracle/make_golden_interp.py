"""TEST INFRASTRUCTURE ONLY.  tests/golden/interp_*.npz: the reference's own `Transformer.make_interpolated_features`
(models/detr/dab_transformer.py:239-294) applied to per-level feature maps, followed by the stack / key-frame slice / rearrange
lines of `Transformer.forward` (:369-392, restated here verbatim in einops terms because they live inside forward()).
Run in the build container only:   python -m oracle.make_golden_interp"""
import os
import numpy as np
import torch
from einops import rearrange

from .ref_import import import_reference
from .make_golden import GOLD

# (fixture, B, shapes, num_frames, eff, seed)
CASES = [
    ("interp_2d_eff", 2, [(4, 6, 6), (4, 3, 3), (4, 2, 2), (4, 1, 1)], 4, True, 0),        # T == num_frames: per-frame 2-D path
    ("interp_2d_all", 1, [(2, 5, 5), (2, 3, 3), (2, 4, 4), (2, 2, 2)], 2, False, 1),
    ("interp_3d_eff", 2, [(4, 6, 5), (4, 3, 3), (2, 3, 4), (2, 2, 1)], 8, True, 2),        # T != num_frames: trilinear path
    ("interp_3d_all", 1, [(3, 4, 4), (3, 2, 2), (2, 5, 3), (1, 2, 2)], 4, False, 3),
]


def main():
    ref = import_reference()
    for name, B, shapes, nf, eff, seed in CASES:
        rs = np.random.RandomState(6000 + seed)
        L = len(shapes)
        Len = sum(t * h * w for t, h, w in shapes)
        tokens = rs.standard_normal((B, Len, 256)).astype(np.float32)
        pos_tokens = rs.standard_normal((B, Len, 256)).astype(np.float32)
        lsi = np.concatenate(([0], np.cumsum([t * h * w for t, h, w in shapes])[:-1])).astype(np.int64)
        feats, poses = [], []
        for l, (t, h, w) in enumerate(shapes):           # :356-365 un-flatten
            sl = slice(int(lsi[l]), int(lsi[l]) + t * h * w)
            feats.append(torch.from_numpy(tokens[:, sl]).reshape(B, t, h, w, 256).permute(0, 4, 1, 2, 3).contiguous())
            poses.append(torch.from_numpy(pos_tokens[:, sl]).reshape(B, t, h, w, 256).permute(0, 4, 1, 2, 3).contiguous())
        inter, pos_l = ref.Transformer.make_interpolated_features(None, feats, poses, level=-2, num_frames=nf)
        srcs = torch.stack(inter, dim=-1)                 # :375  bs, c, t, h, w, l
        pp = torch.stack(pos_l, dim=-1)
        t = srcs.shape[2]
        if eff:                                           # :381-384
            srcs = srcs[:, :, t // 2:t // 2 + 1]
            pp = pp[:, :, t // 2:t // 2 + 1]
        memory = rearrange(srcs, "B C T H W L -> L (H W) (B T) C").contiguous().numpy()      # :391
        pos = rearrange(pp, "B C T H W L -> L (H W) (B T) C").contiguous().numpy()           # :392
        assert np.array_equal(pos[0], pos[-1])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), memory=memory, pos0=pos[0], shapes=np.array(shapes, dtype=np.int64),
                            level_start=lsi, meta=np.array([B, nf, int(eff), seed], dtype=np.int64))
        print(name, memory.shape)


if __name__ == "__main__":
    main()
