"""FusedAdam: torch.optim.Adam semantics (no weight decay / amsgrad) in ONE kernel pass per
parameter (csrc/train_ops.cu), for the 2 GB `_semantic_feature` table (replaces the
Adam built at /root/reference/scene/gaussian_model.py:208 and stepped at train.py:222-223).
State keys (`step`, `exp_avg`, `exp_avg_sq`) match torch.optim.Adam so capture()/restore()
checkpoints interchange."""
from __future__ import annotations

import torch

from . import _C


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False):
        """zero_grad=True also zeroes .grad in the same pass (keeps the buffer allocated, which is
        what lets the backward accumulate into it without a separate memset)."""
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                _C.require_cuda(p, p.grad)
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                _C.check(_C.lib.gags_adam_step(p.data_ptr(), g.data_ptr(),
                                               st["exp_avg"].data_ptr(),
                                               st["exp_avg_sq"].data_ptr(), p.numel(),
                                               float(group["lr"]), float(b1), float(b2),
                                               float(group["eps"]), int(st["step"].item()),
                                               1 if zero_grad else 0, _C.stream_ptr()),
                         "gags_adam_step")
                _C.count_launch()
                if zero_grad and g is not p.grad:
                    p.grad.zero_()
        return loss

    @torch.no_grad()
    def step_chunks(self, p, ready):
        """One Adam step on parameter `p`, issued slice by slice: `ready(i, n)` is called before
        slice i of n is updated (e.g. to wait for that slice of an asynchronous gradient
        all-reduce), so the reduction of slice i+1 overlaps the update of slice i.  Same arithmetic
        as step() — the kernel is elementwise and the step counter advances once."""
        group = next(g for g in self.param_groups if any(q is p for q in g["params"]))
        b1, b2 = group["betas"]
        _C.require_cuda(p, p.grad)
        st = self.state[p]
        if len(st) == 0:
            st["step"] = torch.tensor(0.0)
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        st["step"] += 1
        g = p.grad
        if not (g.is_contiguous() and p.is_contiguous()):
            raise ValueError("step_chunks needs contiguous parameter and gradient")
        n = ready(None, None)
        numel = p.numel()
        per = -(-numel // n)
        per += (-per) % 4                                  # 16-byte aligned slices
        for i in range(n):
            s0, s1 = i * per, min(numel, (i + 1) * per)
            if s0 >= s1:
                break
            ready(i, n)
            _C.check(_C.lib.gags_adam_step(p.data_ptr() + 4 * s0, g.data_ptr() + 4 * s0,
                                           st["exp_avg"].data_ptr() + 4 * s0,
                                           st["exp_avg_sq"].data_ptr() + 4 * s0, s1 - s0,
                                           float(group["lr"]), float(b1), float(b2),
                                           float(group["eps"]), int(st["step"].item()), 0,
                                           _C.stream_ptr()), "gags_adam_step")
            _C.count_launch()
