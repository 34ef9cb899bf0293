"""Thin torch-facing wrappers of the libcqvad building blocks (no autograd; inference path).  Every function here launches
kernels of libcqvad.so for its arithmetic -- there is no torch fallback; dtype casts / .contiguous() of the arguments are torch
tensor plumbing.  (The stand-alone per-layer module forwards that call these keep some small glue in torch: see their docstrings.)"""
import torch

from .. import _lib


def _w(t, dtype):
    return t.detach().to(dtype).contiguous()


def linear(x, weight, bias=None, act=_lib.ACT_NONE, res=None):
    """act(x @ weight.T + bias) (+ res) through cqvad_linear.  x [..., K]; weight [N, K] (or 1x1 conv [N,K,1,1])."""
    _lib.require_cuda(x)
    _lib.warn_if_grad("ops.linear", x, weight, bias, res)
    dt = x.dtype
    K = x.shape[-1]
    x2 = x.reshape(-1, K).contiguous()
    w2 = _w(weight.reshape(weight.shape[0], -1), dt)
    N = w2.shape[0]
    b = None if bias is None else _w(bias, torch.float32)
    r = None if res is None else res.reshape(-1, N).to(dt).contiguous()
    out = torch.empty((x2.shape[0], N), dtype=dt, device=x.device)
    p = _lib.ptr
    _lib.check(_lib.lib().cqvad_linear(_lib.dtype_id(dt), p(x2), p(w2), p(b), p(r), p(out), x2.shape[0], N, K, act,
                                       _lib.stream_ptr()))
    return out.view(*x.shape[:-1], N)


def layer_norm(x, weight, bias, eps=1e-5, res=None):
    """LayerNorm(x (+ res)) over the last dim (256) through cqvad_layernorm."""
    _lib.require_cuda(x)
    _lib.warn_if_grad("ops.layer_norm", x, weight, bias, res)
    dt = x.dtype
    C = x.shape[-1]
    x2 = x.reshape(-1, C).contiguous()
    r = None if res is None else res.reshape(-1, C).to(dt).contiguous()
    out = torch.empty_like(x2)
    p = _lib.ptr
    g, b = _w(weight, torch.float32), _w(bias, torch.float32)   # keep the converted copies alive across the launch
    _lib.check(_lib.lib().cqvad_layernorm(_lib.dtype_id(dt), p(x2), p(r), p(g), p(b), float(eps), p(out), 0,
                                          x2.shape[0], C, _lib.stream_ptr()))
    return out.view(x.shape)


def ffn(x, w1, b1, w2, b2, act=_lib.ACT_RELU, res=None, ln_weight=None, ln_bias=None, eps=1e-5):
    """LN?( res + w2 . act(w1 . x + b1) + b2 ) through cqvad_mlp (fused tcgen05 kernel in bf16 when F % 128 == 0)."""
    _lib.require_cuda(x)
    _lib.warn_if_grad("ops.ffn", x, w1, w2, res)
    dt = x.dtype
    C = x.shape[-1]
    x2 = x.reshape(-1, C).contiguous()
    W1, W2 = _w(w1, dt), _w(w2, dt)
    B1, B2 = _w(b1, torch.float32), _w(b2, torch.float32)
    F = W1.shape[0]
    r = None if res is None else res.reshape(-1, C).to(dt).contiguous()
    g = None if ln_weight is None else _w(ln_weight, torch.float32)
    b = None if ln_bias is None else _w(ln_bias, torch.float32)
    out = torch.empty_like(x2)
    hid = torch.empty((x2.shape[0], F), dtype=dt, device=x.device)
    p = _lib.ptr
    _lib.check(_lib.lib().cqvad_mlp(_lib.dtype_id(dt), p(x2), p(W1), p(B1), p(W2), p(B2), act, p(r), p(g), p(b), float(eps),
                                    p(out), p(hid), x2.shape[0], F, _lib.stream_ptr()))
    return out.view(x.shape)


def mha_core(q, k, v, num_heads, key_padding_mask=None, query_specific_key=False):
    """Attention core (attention.py:336-414) through cqvad_mha_core.  Returns [L, Nb, Ev] before out_proj."""
    _lib.require_cuda(q, k, v)
    _lib.warn_if_grad("ops.mha_core", q, k, v)
    dt = q.dtype
    L, Nb, E = q.shape
    S = k.shape[1] if query_specific_key else k.shape[0]
    Ev = v.shape[-1]
    q, k, v = q.contiguous(), k.to(dt).contiguous(), v.to(dt).contiguous()
    kpm = None if key_padding_mask is None else key_padding_mask.to(torch.uint8).contiguous()
    out = torch.empty((L, Nb, Ev), dtype=dt, device=q.device)
    p = _lib.ptr
    _lib.check(_lib.lib().cqvad_mha_core(_lib.dtype_id(dt), 1 if query_specific_key else 0, p(q), p(k), p(v), p(kpm),
                                         p(out), L, S, Nb, num_heads, E, Ev, _lib.stream_ptr()))
    return out


def conv_block(x_nhwc, conv1_w, conv1_b, ln_w, ln_b, conv2_w, conv2_b, conv3_w, conv3_b):
    """ConvBlock.forward (dab_transformer.py:88-98) on NHWC input through cqvad_convblock_forward."""
    _lib.require_cuda(x_nhwc)
    dt = x_nhwc.dtype
    n, h, w, C = x_nhwc.shape
    x = x_nhwc.contiguous()
    y = torch.empty_like(x)
    w1 = _w(conv1_w.permute(0, 2, 3, 1).reshape(conv1_w.shape[0], -1), dt)
    lib = _lib.lib()
    nbytes = lib.cqvad_convblock_workspace_bytes(_lib.dtype_id(dt), n, h, w)
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
    off = (-ws.data_ptr()) % 1024
    f32 = lambda t: _w(t, torch.float32)
    p = _lib.ptr
    import ctypes
    # converted copies are bound to names so they outlive the (asynchronous) launch
    b1, g, b, w2, b2, w3, b3 = f32(conv1_b), f32(ln_w), f32(ln_b), _w(conv2_w, dt), f32(conv2_b), _w(conv3_w, dt), f32(conv3_b)
    _lib.check(lib.cqvad_convblock_forward(_lib.dtype_id(dt), p(x), p(y), p(w1), p(b1), p(g), p(b), p(w2), p(b2), p(w3),
                                           p(b3), n, h, w, ctypes.c_void_p(ws.data_ptr() + off), nbytes,
                                           _lib.stream_ptr()))
    return y
