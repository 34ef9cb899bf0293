"""Drop-ins for models/position_encoding.py (PositionEmbeddingSine_3D, build_position_encoding) and
dab_transformer.gen_sineembed_for_position, computed by cqvad_posenc3d / cqvad_sine_embed."""
import math

import torch
from torch import nn

from .. import _lib


class PositionEmbeddingSine_3D(nn.Module):
    def __init__(self, num_pos_feats=64, temperature=10000, normalize=False, scale=None):
        super().__init__()
        if scale is not None and normalize is False:
            raise ValueError("normalize should be True if scale is passed")
        if not normalize or temperature != 10000 or (scale is not None and abs(scale - 2 * math.pi) > 1e-12):
            raise NotImplementedError("only the configuration built by build_position_encoding "
                                      "(normalize=True, temperature=10000, scale=2*pi) is implemented")
        self.num_pos_feats = num_pos_feats
        self.num_pos_feats_t = num_pos_feats / 8 * 2
        self.num_pos_feats_s = num_pos_feats / 8 * 3
        self.temperature, self.normalize, self.scale = temperature, normalize, 2 * math.pi

    def forward(self, tensor_list):
        """tensor_list: NestedTensor-like with .tensors (only its device is used) and .mask [B,T,H,W] bool."""
        mask = tensor_list.mask
        assert mask is not None
        _lib.require_cuda(mask)
        B, T, H, W = mask.shape
        m8 = mask.to(torch.uint8).contiguous()
        pos = torch.empty((B, self.num_pos_feats, T, H, W), dtype=torch.float32, device=mask.device)
        _lib.check(_lib.lib().cqvad_posenc3d(_lib.ptr(m8), _lib.ptr(pos), B, T, H, W, self.num_pos_feats, _lib.stream_ptr()))
        return pos


def build_position_encoding(hidden_dim):
    return PositionEmbeddingSine_3D(hidden_dim, normalize=True)


def gen_sineembed_for_position(pos_tensor):
    """[nq, bs, 4] -> [nq, bs, 512] (dab_transformer.py:50-76; only the 4-coordinate form used by the decoder)."""
    if pos_tensor.size(-1) != 4:
        raise ValueError("Unknown pos_tensor shape(-1):{}".format(pos_tensor.size(-1)))
    _lib.require_cuda(pos_tensor)
    p = pos_tensor.float().contiguous()
    rows = p.numel() // 4
    out = torch.empty((*p.shape[:-1], 512), dtype=torch.float32, device=p.device)
    _lib.check(_lib.lib().cqvad_sine_embed(_lib.ptr(p), _lib.ptr(out), rows, _lib.stream_ptr()))
    return out
