"""torch.autograd.Function that puts the native decoder training step behind `loss.backward()`.

The reference has no hand-written backward for the decoder: autograd differentiates TransformerDecoder.forward
(models/detr/dab_transformer.py:722-852; train.py:151).  Here the forward is cqvad_decoder_train_forward and the backward
cqvad_decoder_backward (include/cqvad.h); this Function only routes tensors: inputs (tgt, memory, refpoints_unsigmoid) and the
module's parameters in, their gradients out.  `pos` and the mask get no gradient (see the header)."""
import torch


class DecoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, names, mask, pos, orig_res, tgt, memory, refpoints_unsigmoid, *params):
        p_drop, seed = getattr(engine, "train_dropout", (0.0, 0))
        out = engine.forward_train(tgt.detach(), memory.detach(), mask, pos.detach(), refpoints_unsigmoid.detach(), orig_res,
                                   dropout_p=p_drop, seed=seed)
        ctx.engine, ctx.names, ctx.generation = engine, names, out["generation"]
        ctx.need = (tgt.requires_grad, memory.requires_grad, refpoints_unsigmoid.requires_grad)
        ctx.in_dtypes = (tgt.dtype, memory.dtype, refpoints_unsigmoid.dtype)
        ctx.param_meta = [(p.shape, p.dtype) for p in params]
        ctx.params = params
        ctx.mark_non_differentiable()
        return out["hs"], out["cls_hs"], out["refs"]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_hs, g_cls, g_refs):
        # fast path: every parameter already owns a contiguous fp32 .grad (e.g. the persistent views of FlatAdamW): the native
        # backward accumulates into them in place and autograd receives no per-parameter tensors at all
        grads = {n: p.grad for n, p in zip(ctx.names, ctx.params)}
        used = [n for n in ctx.names if not ("q_proj." in n or n.startswith("cls_norm."))]
        if all(grads[n] is not None and grads[n].dtype == torch.float32 and grads[n].is_contiguous() and grads[n].is_cuda for n in used):
            g = ctx.engine.backward_into({n: grads[n] for n in used}, g_hs, g_cls, g_refs, generation=ctx.generation)
            gt = g["tgt"].to(ctx.in_dtypes[0]).clone() if ctx.need[0] else None
            gm = g["memory"].to(ctx.in_dtypes[1]).clone() if ctx.need[1] else None
            gr = g["refpoints_unsigmoid"].to(ctx.in_dtypes[2]).clone() if ctx.need[2] else None
            return (None, None, None, None, None, gt, gm, gr, *([None] * len(ctx.params)))
        g = ctx.engine.backward(g_hs, g_cls, g_refs, zero=True, named=True, generation=ctx.generation)
        named = g["params"]
        pg = []
        for name, (shape, dtype) in zip(ctx.names, ctx.param_meta):
            t = named.get(name)        # None: parameter the reference never uses (q_proj, cls_norm): grad None like autograd
            pg.append(None if t is None else t.reshape(shape).to(dtype))
        gt = g["tgt"].to(ctx.in_dtypes[0]).clone() if ctx.need[0] else None
        gm = g["memory"].to(ctx.in_dtypes[1]).clone() if ctx.need[1] else None
        gr = g["refpoints_unsigmoid"].to(ctx.in_dtypes[2]).clone() if ctx.need[2] else None
        return (None, None, None, None, None, gt, gm, gr, *pg)
