"""Drop-in for the ViT simple-feature-pyramid neck of the reference Backbone (`lateral_convs`, models/backbone_3d_builder.py:133-182,
applied by `space_forward` :190-200): same sub-module layout and parameter names (`lateral_convs.{level}.{index}.weight|bias`), so
the matching entries of a reference checkpoint load with strict=True.  Each level is ONE C-ABI call (cqvad_vit_neck_level) that
writes straight into the encoder's token sequence; `forward` reshapes that to the reference's channel-first maps for callers
that want `space_forward`'s output format.  Inference path (no backward: the reference trains the neck together with the ViT
backbone, which is outside this library's scope)."""
import ctypes

import torch
from torch import nn

from .. import _lib

SCALES = (4.0, 2.0, 1.0, 0.5)


class ChannelLayerNorm(nn.Module):
    """Parameter container of the reference's channel-first LayerNorm (backbone_3d_builder.py:20-40, eps 1e-6)."""

    def __init__(self, normalized_shape, eps=1e-6):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps


def _lateral(dim, scale, out_channel):
    # the module list of backbone_3d_builder.py:139-180 (parameter containers: the arithmetic runs in libcqvad)
    if scale == 4.0:
        layers = [nn.ConvTranspose3d(dim, dim // 2, kernel_size=[1, 2, 2], stride=[1, 2, 2]), ChannelLayerNorm(dim // 2), nn.GELU(),
                  nn.ConvTranspose3d(dim // 2, dim // 4, kernel_size=[1, 2, 2], stride=[1, 2, 2])]
        out_dim = dim // 4
    elif scale == 2.0:
        layers, out_dim = [nn.ConvTranspose3d(dim, dim // 2, kernel_size=[1, 2, 2], stride=[1, 2, 2])], dim // 2
    elif scale == 1.0:
        layers, out_dim = [], dim
    elif scale == 0.5:
        layers, out_dim = [nn.MaxPool3d(kernel_size=[1, 2, 2], stride=[1, 2, 2])], dim
    else:
        raise NotImplementedError(f"scale_factor={scale} is not supported yet.")
    layers.extend([nn.Conv3d(out_dim, out_channel, kernel_size=1, bias=False), ChannelLayerNorm(out_channel),
                   nn.Conv3d(out_channel, out_channel, kernel_size=3, padding=1, bias=False)])
    return nn.Sequential(*layers)


def pack_neck_weights(seq, kind, dtype, device):
    """-> (keepalive list, ctypes pointer table) in the order csrc/neck.cu documents."""
    f32 = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous()
    mat = lambda t: t.detach().to(device=device, dtype=torch.float32).contiguous().to(dtype).contiguous()

    def convT(m):     # weight [Cin, Cout, 1, 2, 2] -> [(dy, dx, co), ci]; bias tiled over the four (dy, dx) positions
        w = m.weight.detach()[:, :, 0].permute(2, 3, 1, 0).reshape(-1, m.in_channels)
        return [mat(w), f32(m.bias.detach().repeat(4))]

    mods = list(seq)
    keep = []
    i = 0
    if kind == 0:
        keep += convT(mods[0]) + [f32(mods[1].weight), f32(mods[1].bias)] + convT(mods[3])
        i = 4
    elif kind == 1:
        keep += convT(mods[0])
        i = 1
    elif kind == 3:
        i = 1
    c1, ln, c3 = mods[i], mods[i + 1], mods[i + 2]
    keep += [mat(c1.weight.reshape(c1.out_channels, c1.in_channels)), f32(ln.weight), f32(ln.bias),
             mat(c3.weight.detach().permute(2, 0, 3, 4, 1).reshape(3, 256, 9 * 256))]
    tab = (ctypes.c_void_p * len(keep))(*[k.data_ptr() for k in keep])
    return keep, tab


class SimpleFeaturePyramid(nn.Module):
    def __init__(self, embed_dim=768, out_channel=256):
        super().__init__()
        if out_channel != 256:
            raise ValueError("libcqvad neck: D_MODEL 256")
        self.lateral_convs = nn.ModuleList([_lateral(embed_dim, s, out_channel) for s in SCALES])
        self.embed_dim = embed_dim
        self._packed = {}

    def weights_updated(self):
        self._packed = {}

    def _pack(self, kind, dtype, device):
        key = (kind, dtype, str(device), tuple(p._version for p in self.lateral_convs[kind].parameters()))
        if self._packed.get(kind, (None,))[0] != key:
            self._packed[kind] = (key, pack_neck_weights(self.lateral_convs[kind], kind, dtype, device))
        return self._packed[kind][1]

    @torch.no_grad()
    def forward_tokens(self, features):
        """features: the four ViT feature maps [B, C, T, H, W] (one per level).  Returns (src_flatten [B, Len, 256], shapes [4, 3],
        level_start [4]) -- the encoder's input, without materialising channel-first maps."""
        _lib.require_cuda(*features)
        lib = _lib.lib()
        dt, dev = features[0].dtype, features[0].device
        B, C, T, H, W = features[0].shape
        shapes = [(T, 4 * H, 4 * W), (T, 2 * H, 2 * W), (T, H, W), (T, H // 2, W // 2)]
        ns = [t * h * w for t, h, w in shapes]
        Len = sum(ns)
        tokens = torch.empty((B, Len, 256), dtype=dt, device=dev)
        start = 0
        p = _lib.ptr
        for kind, f in enumerate(features):
            if tuple(f.shape) != (B, C, T, H, W):
                raise ValueError("the four ViT features share one shape")
            keep, tab = self._pack(kind, dt, dev)
            need = lib.cqvad_vit_neck_workspace_bytes(_lib.dtype_id(dt), kind, B, C, T, H, W)
            if need == 0:
                _lib.check(-1)
            ws = torch.empty(need, dtype=torch.uint8, device=dev)
            fc = f.contiguous()
            _lib.check(lib.cqvad_vit_neck_level(_lib.dtype_id(dt), kind, p(fc), tab, p(tokens), Len, start, p(ws), need, B, C, T, H, W,
                                                _lib.stream_ptr()))
            start += ns[kind]
        sh, ls = _lib.shape_tensors(shapes, dev)
        return tokens, sh, ls

    @torch.no_grad()
    def forward(self, features):
        """`space_forward` format: {"0": [B,256,T,4H,4W], "1": ..., "2": ..., "3": ...} (a re-layout of forward_tokens' result)."""
        tokens, sh, ls = self.forward_tokens(features)
        out = {}
        start = 0
        for l, (t, h, w) in enumerate(_lib.host_shapes(sh)):
            s, start = start, start + t * h * w
            out[str(l)] = tokens[:, s:s + t * h * w].reshape(tokens.shape[0], t, h, w, 256).permute(0, 4, 1, 2, 3).contiguous()
        return out
