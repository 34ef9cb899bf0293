"""Host driver of the native decoder: packs a reference-named state_dict into the weight table of
cqvad_decoder_forward (include/cqvad.h) and owns the device workspace.  torch is used for device memory and streams
only; every arithmetic step is a kernel of libcqvad.so."""
import ctypes
from ctypes import c_void_p, byref

import numpy as np
import torch

from . import _lib


def _as_tensor(v):
    if isinstance(v, torch.Tensor):
        return v.detach()
    return torch.from_numpy(np.ascontiguousarray(v))


def pack_decoder_weights(state, layers, dtype, device, meta=None):
    """state: mapping of reference TransformerDecoder state_dict names (+ optional
    'heads.class_embed_b.*') to tensors / ndarrays.  Returns (keepalive list, ctypes pointer array).

    Matrices are stored in `dtype`; biases, LayerNorm parameters and the four small-N linears stay fp32; the 3x3 conv
    weight [O,I,3,3] is re-laid out as [O][ky*3+kx][I] (K-major for the implicit GEMM); 1x1 conv weights are [O,I]."""
    lib = _lib.lib()
    n = lib.cqvad_decoder_num_weights(layers)
    keep, arr = [], (c_void_p * n)()
    if meta is not None:
        meta.clear()
    for i in range(n):
        name = lib.cqvad_decoder_weight_name(i, layers).decode()
        kind = lib.cqvad_decoder_weight_kind(i, layers)
        if ".__ca_kv." in name:   # synthesised: [ca_kcontent_proj ; ca_v_proj] stacked along the output dimension
            leaf = name.rsplit(".", 1)[1]
            pre = name.split(".__ca_kv.")[0]
            t = torch.cat([_as_tensor(state[f"{pre}.ca_kcontent_proj.{leaf}"]).float(), _as_tensor(state[f"{pre}.ca_v_proj.{leaf}"]).float()], 0)
            t = t.to(device=device).contiguous().to(torch.float32 if kind == 1 else dtype).contiguous()
            keep.append(t)
            arr[i] = t.data_ptr()
            if meta is not None:
                meta.append((name, tuple(t.shape), None))
            continue
        if name not in state:
            if "ca_qpos_proj" in name and not name.startswith("layers.0."):
                arr[i] = None   # ca_qpos_proj is None for layers >= 1 (dab_transformer.py:711-713)
                keep.append(None)
                if meta is not None:
                    meta.append((name, None, None))
                continue
            if name.startswith("heads.class_embed_b"):
                t = torch.zeros((3, 256) if name.endswith("weight") else (3,))
            else:
                raise KeyError(f"decoder weight '{name}' missing from the state dict")
        else:
            t = _as_tensor(state[name])
        t = t.to(device=device, dtype=torch.float32)
        orig_shape = tuple(t.shape)
        if name.endswith("conv1.weight"):
            t = t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)           # [O,I,3,3] -> [O, (ky,kx,I)]
        elif t.dim() == 4:
            t = t.reshape(t.shape[0], t.shape[1])                        # 1x1 conv
        t = t.contiguous().to(torch.float32 if kind == 1 else dtype).contiguous()
        keep.append(t)
        arr[i] = t.data_ptr()
        if meta is not None:
            meta.append((name, tuple(t.shape), orig_shape))
    return keep, arr


def grad_bucket_layout(names, sizes, layers):
    """Offsets of every weight gradient in the flat fp32 buffer and the bucket boundaries: one contiguous bucket per decoder layer
    (layers.l.* then cls_layers.l.*), then the shared modules.  Returns (offs [n+1], [(lo, hi)] * (layers + 1))."""
    n = len(sizes)
    layer_of = []
    for name in names:
        parts = name.split(".")
        layer_of.append(int(parts[1]) if parts[0] in ("layers", "cls_layers") and parts[1].isdigit() else layers)
    order = sorted(range(n), key=lambda i: (layer_of[i], i))
    offs = np.zeros(n + 1, dtype=np.int64)
    pos, bounds = 0, [0] * (layers + 2)
    for i in order:
        offs[i] = pos
        pos += (sizes[i] + 63) // 64 * 64
        bounds[layer_of[i] + 1] = pos
    for b in range(1, len(bounds)):
        bounds[b] = max(bounds[b], bounds[b - 1])
    offs[n] = pos
    return offs, [(bounds[b], bounds[b + 1]) for b in range(layers + 1)]


class GraphedForward:
    """A captured DecoderEngine.forward (DecoderEngine.capture_forward)."""

    def __init__(self, graph, inputs, outputs):
        self.graph, self.inputs, self.outputs = graph, inputs, outputs

    def __call__(self, **new_inputs):
        for k, v in new_inputs.items():
            self.inputs[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.outputs


class DecoderEngine:
    """TransformerDecoder.forward (+ DETR heads) on one GPU.

    >>> eng = DecoderEngine(state_dict, nq=15, K=80, layers=6, F=2048, dtype=torch.bfloat16, device="cuda")
    >>> out = eng.forward(tgt, memory, mask, pos, refpoints_unsigmoid, (h, w))
    """

    def __init__(self, state, nq, K, layers, F=2048, dtype=torch.bfloat16, device="cuda", out_f32=True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("Not implemented on the CPU")
        self.dtype, self.nq, self.K, self.layers, self.F, self.out_f32 = dtype, nq, K, layers, F, out_f32
        self._meta = []
        self._keep, self._wtab = pack_decoder_weights(state, layers, dtype, self.device, self._meta)
        self._ws = None
        self._tws = None
        self._train_ctx = None
        self._train_gen = 0          # generation of the saved activations in _tws (one live training forward per engine)
        self.last_launches = 0

    def repack(self, state):
        """Refresh the packed weights IN PLACE after an optimizer step: pointer table, workspaces and the flat gradient buffer stay."""
        plain_dst, plain_src = [], []
        for (name, shp, orig), old in zip(self._meta, self._keep):
            if old is None:
                continue
            if ".__ca_kv." in name:
                pre, leaf = name.split(".__ca_kv.")
                half = old.shape[0] // 2
                plain_dst += [old[:half], old[half:]]
                plain_src += [_as_tensor(state[f"{pre}.ca_kcontent_proj.{leaf}"]), _as_tensor(state[f"{pre}.ca_v_proj.{leaf}"])]
            elif name not in state:
                continue                                   # synthesised zeros (absent heads): nothing to refresh
            elif name.endswith("conv1.weight"):
                src = _as_tensor(state[name])
                old.view(src.shape[0], 3, 3, src.shape[1]).copy_(src.permute(0, 2, 3, 1))
            else:
                plain_dst.append(old)
                plain_src.append(_as_tensor(state[name]).reshape(old.shape))
        if plain_dst:
            torch._foreach_copy_(plain_dst, plain_src)      # one multi-tensor cast / copy launch group instead of ~4 kernels per weight

    def _workspace(self, desc):
        need = _lib.lib().cqvad_decoder_workspace_bytes(byref(desc))
        if need == 0:
            _lib.check(-1)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, tgt, memory, mask, pos, refpoints_unsigmoid, orig_res, heads=True, skip_cls_hs=False, fp32_cls_stream=False):
        """tgt [nq,BT,256], memory/pos [4,S,BT,256], mask [BT,S] bool or None, refpoints_unsigmoid [nq,BT,4]
        (layouts of dab_transformer.py:391-396).  Returns dict(hs, cls_hs, refs[, pred_logits, pred_boxes, pred_logits_b]).
        fp32_cls_stream (bf16 path): CQVAD_DEC_FP32_CLS_STREAM -- fp32 side copies of the class-token residual / output stream
        (+0.33 ms per 32-clip step; measured to leave the bf16 error unchanged, which is dominated by GEMM operand rounding)."""
        lib = _lib.lib()
        h, w = orig_res
        nq, BT = tgt.shape[0], tgt.shape[1]
        S = h * w
        if nq != self.nq or tuple(memory.shape) != (4, S, BT, 256) or tuple(refpoints_unsigmoid.shape) != (nq, BT, 4):
            raise ValueError(f"decoder input shapes do not match: tgt {tuple(tgt.shape)}, memory {tuple(memory.shape)}, "
                             f"refpoints {tuple(refpoints_unsigmoid.shape)}, orig_res {orig_res}")
        _lib.require_cuda(tgt, memory, pos, refpoints_unsigmoid)
        f32 = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        tgt, memory, pos0, ref = f32(tgt), f32(memory), f32(pos[0]), f32(refpoints_unsigmoid)
        m8 = None
        if mask is not None:
            m8 = mask.to(self.device).contiguous()
            m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else m8.to(torch.uint8)   # zero-copy for bool masks
        desc = _lib.DecoderDesc(_lib.dtype_id(self.dtype), BT, nq, h, w, self.K, self.F, self.layers,
                                1 if self.out_f32 else 0,
                                (_lib.DEC_SKIP_CLS_HS if skip_cls_hs else 0) | (_lib.DEC_FP32_CLS_STREAM if fp32_cls_stream else 0))
        ws = self._workspace(desc)
        odt = torch.float32 if self.out_f32 else self.dtype
        Lr, K = self.layers, self.K
        hs = torch.empty((Lr, BT, nq, 256), dtype=odt, device=self.device)
        cls_hs = None if skip_cls_hs else torch.empty((Lr, BT, nq, K, 256), dtype=odt, device=self.device)
        refs = torch.empty((Lr, BT, nq, 4), dtype=torch.float32, device=self.device)
        pl = pb = plb = None
        if heads:
            pl = torch.empty((Lr, BT, nq, K), dtype=torch.float32, device=self.device)
            pb = torch.empty((Lr, BT, nq, 4), dtype=torch.float32, device=self.device)
            plb = torch.empty((Lr, BT, nq, 3), dtype=torch.float32, device=self.device)
        p = _lib.ptr
        rc = lib.cqvad_decoder_forward(byref(desc), self._wtab, p(tgt), p(memory), p(pos0), p(m8), p(ref), p(hs), p(cls_hs),
                                       p(refs), p(pl), p(pb), p(plb), p(ws), ws.numel(), _lib.stream_ptr())
        _lib.check(rc)
        self.last_launches = lib.cqvad_last_launch_count()
        out = dict(hs=hs, cls_hs=cls_hs, refs=refs)
        if heads:
            out.update(pred_logits=pl, pred_boxes=pb, pred_logits_b=plb)
        return out

    def capture_forward(self, tgt, memory, mask, pos, refpoints_unsigmoid, orig_res, heads=True, skip_cls_hs=False):
        """CUDA-graph capture of one inference forward at fixed shapes (launch-bound at small batch: 317 kernels on two streams).
        Returns a GraphedForward: `g.inputs` are the static input tensors (copy new data into them), `g()` replays the graph on
        the current stream and returns the static output dict.  The library's side stream joins the capture through its
        fork / join events; tensor maps are kernel parameters, so nothing is re-encoded on replay."""
        cur = torch.cuda.current_stream()
        f32 = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous().clone()
        static = dict(tgt=f32(tgt), memory=f32(memory), pos=f32(pos), refpoints_unsigmoid=f32(refpoints_unsigmoid),
                      mask=None if mask is None else mask.to(self.device).contiguous().clone())
        run = lambda: self.forward(static["tgt"], static["memory"], static["mask"], static["pos"], static["refpoints_unsigmoid"], orig_res,
                                   heads=heads, skip_cls_hs=skip_cls_hs)
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(cur)
        with torch.cuda.stream(warm):                 # workspace allocation, kernel attributes, lazy inits: outside the capture
            for _ in range(2):
                run()
        cur.wait_stream(warm)
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = run()
        return GraphedForward(graph, static, out)

    # ---- training step (BASELINE.json configs[1]: decoder fwd + bwd) ----------------------------------------------
    def _grad_table(self):
        """One flat fp32 buffer holding every weight gradient + the pointer table cqvad_decoder_backward takes.  Layout = one
        contiguous bucket per decoder layer (layers.l.* followed by cls_layers.l.*), then the shared modules: the buckets are
        what `dist.allreduce_gradients(overlap=True)` reduces while the backward of the lower layers still runs."""
        if getattr(self, "_gflat", None) is None:
            sizes = [0 if shp is None else int(np.prod(shp)) for (_, shp, _) in self._meta]
            offs, self.grad_buckets = grad_bucket_layout([m[0] for m in self._meta], sizes, self.layers)   # [layer 0 .. L-1, shared]
            pos = offs[-1]
            self._gflat = torch.zeros(int(pos), dtype=torch.float32, device=self.device)
            self._goffs, self._gsizes = offs, sizes
            self._gtab = (c_void_p * len(sizes))()
            for i, sz in enumerate(sizes):
                self._gtab[i] = None if self._meta[i][1] is None else self._gflat.data_ptr() + 4 * int(offs[i])
        return self._gflat, self._gtab

    def enable_layer_events(self, on=True):
        """Have every later backward record one CUDA event per decoder layer as soon as that layer's parameter gradients are final
        (cqvad_decoder_backward_layer_events); `self.layer_events[l]` are torch events a communication stream can wait on."""
        lib = _lib.lib()
        if not on:
            _lib.check(lib.cqvad_decoder_backward_layer_events(None, 0))
            self.layer_events = None
            return
        self.layer_events = [torch.cuda.Event() for _ in range(self.layers)]
        for e in self.layer_events:
            e.record()                                   # creates the underlying cudaEvent_t
        self._evtab = (c_void_p * self.layers)(*[e.cuda_event for e in self.layer_events])
        _lib.check(lib.cqvad_decoder_backward_layer_events(self._evtab, self.layers))

    def forward_train(self, tgt, memory, mask, pos, refpoints_unsigmoid, orig_res, dropout_p=0.0, seed=0):
        """TransformerDecoder.forward keeping what the backward needs.  Returns dict(hs, cls_hs, refs); follow with
        `backward(grad_hs, grad_cls_hs, grad_refs)`.  dropout_p > 0: nn.Dropout at the residual-branch / FFN-hidden sites of every
        layer (dab_transformer.py:937,991,995,1043,1062,1076) with Philox masks keyed by `seed`; the backward regenerates them."""
        lib = _lib.lib()
        h, w = orig_res
        nq, BT = tgt.shape[0], tgt.shape[1]
        S = h * w
        if nq != self.nq or tuple(memory.shape) != (4, S, BT, 256) or tuple(refpoints_unsigmoid.shape) != (nq, BT, 4):
            raise ValueError("decoder input shapes do not match")
        _lib.require_cuda(tgt, memory, pos, refpoints_unsigmoid)
        f32 = lambda t: t.to(device=self.device, dtype=torch.float32).contiguous()
        tgt, memory, pos0, ref = f32(tgt), f32(memory), f32(pos[0]), f32(refpoints_unsigmoid)
        m8 = None
        if mask is not None:
            m8 = mask.to(self.device).contiguous()
            m8 = m8.view(torch.uint8) if m8.dtype == torch.bool else m8.to(torch.uint8)
        desc = _lib.DecoderDesc(_lib.dtype_id(self.dtype), BT, nq, h, w, self.K, self.F, self.layers,
                                1 if self.out_f32 else 0, 0, float(dropout_p), int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
        need = lib.cqvad_decoder_train_workspace_bytes(byref(desc))
        if need == 0:
            _lib.check(-1)
        if self._tws is None or self._tws.numel() < need:
            self._tws = None
            self._tws = torch.empty(need, dtype=torch.uint8, device=self.device)
        odt = torch.float32 if self.out_f32 else self.dtype
        Lr, K = self.layers, self.K
        hs = torch.empty((Lr, BT, nq, 256), dtype=odt, device=self.device)
        cls_hs = torch.empty((Lr, BT, nq, K, 256), dtype=odt, device=self.device)
        refs = torch.empty((Lr, BT, nq, 4), dtype=torch.float32, device=self.device)
        p = _lib.ptr
        rc = lib.cqvad_decoder_train_forward(byref(desc), self._wtab, p(tgt), p(memory), p(pos0), p(m8), p(ref), p(hs),
                                             p(cls_hs), p(refs), p(self._tws), self._tws.numel(), _lib.stream_ptr())
        _lib.check(rc)
        self.last_launches = lib.cqvad_last_launch_count()
        self._train_ctx = (desc, m8, (nq, BT, S))
        self._train_gen += 1
        return dict(hs=hs, cls_hs=cls_hs, refs=refs, generation=self._train_gen)

    def backward(self, grad_hs=None, grad_cls_hs=None, grad_refs=None, zero=True, named=True, generation=None):
        """Backward of the last forward_train.  Returns dict(memory, tgt, refpoints_unsigmoid, params={reference name: grad}).
        `generation` (the value forward_train returned): raises if another forward_train has overwritten the saved activations."""
        if self._train_ctx is None:
            raise RuntimeError("backward() without a preceding forward_train()")
        if generation is not None and generation != self._train_gen:
            raise RuntimeError("DecoderEngine.backward: the activations of this forward were overwritten by a later forward_train() on the "
                               "same engine (one live training forward per engine: run backward before the next forward, or use a "
                               "second TransformerDecoder / DecoderEngine)")
        lib = _lib.lib()
        desc, m8, (nq, BT, S) = self._train_ctx
        odt = torch.float32 if self.out_f32 else self.dtype
        prep = lambda g, dt: None if g is None else g.to(device=self.device, dtype=dt).contiguous()
        grad_hs, grad_cls_hs, grad_refs = prep(grad_hs, odt), prep(grad_cls_hs, odt), prep(grad_refs, torch.float32)
        gflat, gtab = self._grad_table()
        if getattr(self, "_gin", None) is None or self._gin[0].shape[2] != BT or self._gin[0].shape[1] != S:
            self._gin = (torch.zeros((4, S, BT, 256), dtype=torch.float32, device=self.device),
                         torch.zeros((nq, BT, 256), dtype=torch.float32, device=self.device),
                         torch.zeros((nq, BT, 4), dtype=torch.float32, device=self.device))
        elif zero:                       # persistent buffers, zero-filled in place (no allocator traffic in the step)
            for t in self._gin:
                t.zero_()
        if zero:
            gflat.zero_()
        gmem, gtgt, gref = self._gin
        p = _lib.ptr
        rc = lib.cqvad_decoder_backward(byref(desc), self._wtab, p(m8), p(grad_hs), p(grad_cls_hs), p(grad_refs), gtab, p(gmem),
                                        p(gtgt), p(gref), p(self._tws), self._tws.numel(), _lib.stream_ptr())
        _lib.check(rc)
        self.last_launches_bwd = lib.cqvad_last_launch_count()
        out = dict(memory=gmem, tgt=gtgt, refpoints_unsigmoid=gref)
        if named:
            out["params"] = self.named_grads()
        return out

    def backward_into(self, param_grads, grad_hs=None, grad_cls_hs=None, grad_refs=None, generation=None):
        """Backward of the last forward_train that ACCUMULATES the weight gradients straight into caller-owned fp32 tensors
        (`param_grads`: reference parameter name -> contiguous fp32 tensor of the parameter's shape, e.g. the `.grad` views of a
        FlatAdamW buffer): cqvad_decoder_backward adds into them in place, which is autograd's accumulate semantics, so no
        per-parameter clone / add pass exists.  Only the two re-laid-out kinds (3x3 conv taps, stacked [k ; v] projection) go
        through a staging slot of the engine's own buffer.  Returns dict(memory, tgt, refpoints_unsigmoid)."""
        if self._train_ctx is None:
            raise RuntimeError("backward() without a preceding forward_train()")
        if generation is not None and generation != self._train_gen:
            raise RuntimeError("DecoderEngine.backward: the activations of this forward were overwritten by a later forward_train() on the "
                               "same engine (one live training forward per engine)")
        lib = _lib.lib()
        desc, m8, (nq, BT, S) = self._train_ctx
        gflat, gtab0 = self._grad_table()
        key = tuple(t.data_ptr() for t in param_grads.values())
        plan = getattr(self, "_direct_plan", None)
        if plan is None or plan[0] != key:
            tab = (c_void_p * len(self._meta))()
            staged = []                                   # (slot view, [(target tensor, view of the slot in the target's layout)])
            for i, (name, shp, orig) in enumerate(self._meta):
                tab[i] = gtab0[i]
                if shp is None or name.startswith("heads."):
                    continue
                slot = gflat[int(self._goffs[i]):int(self._goffs[i]) + self._gsizes[i]].view(shp)
                if ".__ca_kv." in name:
                    pre, leaf = name.split(".__ca_kv.")
                    half = shp[0] // 2
                    staged.append((slot, [(param_grads[f"{pre}.ca_kcontent_proj.{leaf}"], slot[:half]),
                                          (param_grads[f"{pre}.ca_v_proj.{leaf}"], slot[half:])]))
                elif name.endswith("conv1.weight"):
                    staged.append((slot, [(param_grads[name], slot.view(orig[0], 3, 3, orig[1]).permute(0, 3, 1, 2))]))
                else:
                    tgt = param_grads.get(name)
                    if tgt is None:
                        raise KeyError(f"backward_into: no gradient tensor for '{name}'")
                    if tgt.dtype != torch.float32 or not tgt.is_contiguous() or tgt.numel() != self._gsizes[i]:
                        raise ValueError(f"backward_into: gradient of '{name}' must be a contiguous fp32 tensor of {self._gsizes[i]} elements")
                    tab[i] = tgt.data_ptr()
            plan = (key, tab, staged)
            self._direct_plan = plan
        _, tab, staged = plan
        odt = torch.float32 if self.out_f32 else self.dtype
        prep = lambda g, dt: None if g is None else g.to(device=self.device, dtype=dt).contiguous()
        grad_hs, grad_cls_hs, grad_refs = prep(grad_hs, odt), prep(grad_cls_hs, odt), prep(grad_refs, torch.float32)
        for slot, _ in staged:
            slot.zero_()
        if getattr(self, "_gin", None) is None or self._gin[0].shape[2] != BT:
            self._gin = (torch.empty((4, S, BT, 256), dtype=torch.float32, device=self.device),
                         torch.empty((nq, BT, 256), dtype=torch.float32, device=self.device),
                         torch.empty((nq, BT, 4), dtype=torch.float32, device=self.device))
        gmem, gtgt, gref = self._gin
        for t in self._gin:
            t.zero_()
        p = _lib.ptr
        _lib.check(lib.cqvad_decoder_backward(byref(desc), self._wtab, p(m8), p(grad_hs), p(grad_cls_hs), p(grad_refs), tab, p(gmem),
                                              p(gtgt), p(gref), p(self._tws), self._tws.numel(), _lib.stream_ptr()))
        self.last_launches_bwd = lib.cqvad_last_launch_count()
        for _, targets in staged:
            for tgt, view in targets:
                tgt.add_(view)
        return dict(memory=gmem, tgt=gtgt, refpoints_unsigmoid=gref)

    def named_grads(self):
        """Weight gradients under the reference state_dict names and shapes."""
        res = {}
        for i, (name, shp, orig) in enumerate(self._meta):
            if shp is None or name.startswith("heads."):
                continue
            g = self._gflat[int(self._goffs[i]):int(self._goffs[i]) + self._gsizes[i]].view(shp)
            if ".__ca_kv." in name:
                pre, leaf = name.split(".__ca_kv.")
                half = shp[0] // 2
                for nm, part in ((f"{pre}.ca_kcontent_proj.{leaf}", g[:half]), (f"{pre}.ca_v_proj.{leaf}", g[half:])):
                    res[nm] = res[nm] + part if nm in res else part.clone()
                continue
            if name.endswith("conv1.weight"):
                g = g.view(orig[0], 3, 3, orig[1]).permute(0, 3, 1, 2)
            elif orig is not None and len(orig) == 4:
                g = g.view(orig)
            res[name] = res[name] + g if name in res else g.clone()
        return res
