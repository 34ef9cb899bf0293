"""Optimizer step of the reference training loop (train.py:83,158-167) on ONE flat fp32 buffer:
`torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)` + `torch.optim.AdamW(lr).step()` + `optimizer.zero_grad()` in two
kernel launches (cqvad_adamw_clip_step), and the data-parallel gradient exchange as one NCCL all-reduce over the same buffer
(the reference wraps the model in DDP, utils/model_utils.py:113-121).

Parameters are re-pointed into the flat buffer (`p.data` becomes a view) and `p.grad` is a persistent view of the flat gradient
buffer, so autograd accumulates straight into it and nothing is gathered or scattered per step."""
import re

import torch
import torch.distributed as dist

from . import _lib

# parameters the reference never uses in forward (grad stays None there, so torch's AdamW skips them -- no update, no weight decay)
UNUSED_BY_REFERENCE = re.compile(r"(^|\.)cls_layers\.\d+\.q_proj\.|(^|\.)decoder\.cls_norm\.|^cls_norm\.")


class FlatAdamW:
    def __init__(self, named_params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_norm=1.0, exclude=UNUSED_BY_REFERENCE,
                 module=None):
        """module: the nn.Module that owns the parameters; its submodules exposing `weights_updated()` (packed bf16 weight caches of
        the drop-in decoder / encoder layers) are notified after every step, because the kernel updates the parameters through raw
        pointers and autograd's version counters do not see it."""
        self._listeners = [] if module is None else [m for m in module.modules() if hasattr(m, "weights_updated")]
        named = [(n, p) for n, p in named_params if p.requires_grad and not (exclude is not None and exclude.search(n))]
        if not named:
            raise ValueError("FlatAdamW: no parameters")
        dev = named[0][1].device
        if dev.type != "cuda":
            raise RuntimeError("Not implemented on the CPU")
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        offs, total = [], 0
        for p in self.params:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # 16-byte aligned slots
        self.offsets, self.numel = offs, total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(self.params, offs):
            view = self.flat[o:o + p.numel()].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[o:o + p.numel()].view(p.shape)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self._ws = torch.empty(_lib.lib().cqvad_adamw_workspace_bytes(), dtype=torch.uint8, device=dev)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def zero_grad(self):
        self.grad.zero_()

    def allreduce(self, group=None, async_op=False):
        """SUM all-reduce of the flat gradient; the division by the world size is folded into step(grad_scale=1/world)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def step(self, grad_scale=1.0, zero_grad=True):
        """clip_grad_norm_ + AdamW.step (+ zero_grad) -- parameters change in place (their `_version` is bumped so that packed
        bf16 copies are refreshed)."""
        self.step_count += 1
        p = _lib.ptr
        b1, b2 = self.betas
        _lib.check(_lib.lib().cqvad_adamw_clip_step(p(self.flat), p(self.grad), p(self.exp_avg), p(self.exp_avg_sq), None, self.numel,
                                                    float(self.lr), float(b1), float(b2), float(self.eps), float(self.weight_decay),
                                                    self.step_count, float(self.max_norm), float(grad_scale), 1 if zero_grad else 0,
                                                    p(self.grad_norm), p(self._ws), self._ws.numel(), _lib.stream_ptr()))
        for m in self._listeners:
            m.weights_updated()
        return self.grad_norm
