"""FusedAdam: torch.optim.Adam semantics (no weight decay / amsgrad) in ONE kernel pass per
parameter (csrc/train_ops.cu), for the 2 GB `_semantic_feature` table (replaces the
Adam built at /root/reference/scene/gaussian_model.py:208 and stepped at train.py:222-223).
State keys (`step`, `exp_avg`, `exp_avg_sq`) match torch.optim.Adam so capture()/restore()
checkpoints interchange.

`sparse_rows=True` (what GaussianModel.training_setup(fused_optimizer=True) asks for): a view's
backward touches a small part of a [N, D] table's rows (7 % at BASELINE config 3), yet the dense
pass reads the whole gradient and the training loop re-zeroes it every step.
The optimiser then keeps ONE persistent, always-allocated `.grad` buffer per such parameter plus a
uint8 flag per row; the feature backward reduces straight into the buffer and flags the rows it
can have touched (rasterization.row_flags); step() reads the gradient of flagged rows only, gives
every other row the g = 0 update (bit-identical to the dense pass: m and v decay, the parameter
moves by its momentum) and re-zeroes exactly the flagged rows.  `zero_grad()` keeps the buffer (it
is already zero after step()), i.e. `.grad` is a zero tensor rather than None between steps.  Any
gradient that reaches `.grad` through autograd's own accumulation flags every row (dense
behaviour), and a `.grad` replaced from outside is simply adopted at the next step()."""
from __future__ import annotations

import torch

from . import _C


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, sparse_rows: bool = False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.sparse_rows = bool(sparse_rows)
        self._rows = {}                     # id(param) -> (persistent grad buffer, RowFlags)

    # ---- row-sparse gradient bookkeeping ---------------------------------------------------------
    def _sparse_ok(self, p) -> bool:
        g = p.grad
        return (self.sparse_rows and p.dim() == 2 and p.shape[1] % 4 == 0 and p.is_contiguous()
                and g is not None and g.is_contiguous() and g.dtype == torch.float32
                and p.dtype == torch.float32 and g.shape == p.shape)

    def _adopt(self, p) -> None:
        """Make the (all-zero) `.grad` buffer of `p` the persistent one: row flags, direct reduction
        by the feature backward, and a hook that flags every row when autograd accumulates."""
        from . import rasterization as R
        self._release(p)
        rf = R.RowFlags(torch.zeros(p.shape[0], dtype=torch.uint8, device=p.device))
        self._rows[id(p)] = (p.grad, rf)
        R.row_flags[p.grad.data_ptr()] = rf
        p._gags_direct_grad = True
        if not getattr(p, "_gags_rows_hook", False):
            p._gags_rows_hook = True
            p.register_post_accumulate_grad_hook(self._on_autograd_accumulate)

    def _release(self, p) -> None:
        from . import rasterization as R
        old = self._rows.pop(id(p), None)
        if old is not None:
            R.row_flags.pop(old[0].data_ptr(), None)

    def _on_autograd_accumulate(self, p) -> None:
        rows = self._rows.get(id(p))
        if rows is not None and p.grad is not None and p.grad.data_ptr() == rows[0].data_ptr():
            rows[1].flags.fill_(1)            # a dense gradient arrived through autograd
            rows[1].dirty = True

    def __del__(self):
        try:
            from . import rasterization as R
            for buf, _ in self._rows.values():
                R.row_flags.pop(buf.data_ptr(), None)
        except Exception:
            pass

    def zero_grad(self, set_to_none: bool = True):
        """Persistent row-sparse gradients stay allocated (step() has already re-zeroed them; a
        gradient that was never applied is zeroed here); everything else as torch does."""
        keep = []
        for group in self.param_groups:
            for p in group["params"]:
                rows = self._rows.get(id(p))
                if rows is not None and p.grad is not None \
                        and p.grad.data_ptr() == rows[0].data_ptr():
                    if rows[1].dirty:
                        p.grad.zero_()
                        rows[1].flags.zero_()
                        rows[1].dirty = False
                    keep.append((p, p.grad))
        super().zero_grad(set_to_none=set_to_none)
        for p, g in keep:
            p.grad = g

    @torch.no_grad()
    def step(self, closure=None, zero_grad: bool = False):
        """zero_grad=True also zeroes .grad in the same pass (keeps the buffer allocated, which is
        what lets the backward accumulate into it without a separate memset).  Row-sparse
        parameters always do."""
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
                sparse = self._sparse_ok(p)
                rows = self._rows.get(id(p)) if sparse else None
                if rows is not None and rows[0].data_ptr() == p.grad.data_ptr():
                    _C.check(_C.lib.gags_adam_step_rows(
                        p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                        st["exp_avg_sq"].data_ptr(), rows[1].flags.data_ptr(), p.shape[0],
                        p.shape[1], float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                        int(st["step"].item()), _C.stream_ptr()), "gags_adam_step_rows")
                    _C.count_launch()
                    rows[1].dirty = False
                    continue
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                zg = zero_grad or sparse
                _C.check(_C.lib.gags_adam_step(p.data_ptr(), g.data_ptr(),
                                               st["exp_avg"].data_ptr(),
                                               st["exp_avg_sq"].data_ptr(), p.numel(),
                                               float(group["lr"]), float(b1), float(b2),
                                               float(group["eps"]), int(st["step"].item()),
                                               1 if zg else 0, _C.stream_ptr()),
                         "gags_adam_step")
                _C.count_launch()
                if zg and g is not p.grad:
                    p.grad.zero_()
                if sparse:
                    self._adopt(p)            # this (now all-zero) buffer becomes the persistent one
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
