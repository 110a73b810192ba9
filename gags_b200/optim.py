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
behaviour), and a `.grad` replaced from outside is simply adopted at the next step().

`lazy_rows=True` goes one step further (class LazyRows): the g = 0 update of an untouched row is
not even taken at that step — it is taken later, in registers, when the row is next needed (a view
is about to read it, a gradient arrives for it, or flush()), k steps in one visit with the same
fp32 operations in the same order.  The optimiser pass then touches only the rows the view
touched; the price is that between flushes the parameter tensor holds each row as of ITS last
visit, so anything that reads the table other than render() must call flush() first."""
from __future__ import annotations

import torch

from . import _C


class LazyRows:
    """Book-keeping of the lazily evaluated row-sparse Adam (csrc/train_ops.cu
    adam_lazy_rows_kernel) for ONE [N, D] parameter: `last` (the optimiser step each row is current
    to) and the device table of per-step scalars.  Used by FusedAdam(lazy_rows=True) and
    parallel.SparsePeerAdam(lazy=True).  While rows are behind, the parameter tensor holds, for each
    row, its value as of step last[row]; render() brings the rows a view reads up to date before
    reading them (rasterization.lazy_owners) and flush() materialises the whole table — call it
    before reading the parameter or the moments any other way."""

    def __init__(self, param, exp_avg, exp_avg_sq, betas, eps):
        from . import rasterization as R
        self.param, self.exp_avg, self.exp_avg_sq = param, exp_avg, exp_avg_sq
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        dev = param.device
        self.rows, self.dim = int(param.shape[0]), int(param.shape[1])
        self.last = torch.zeros(self.rows, dtype=torch.int32, device=dev)
        self.t = 0                                   # steps recorded so far
        self.cap = 4096
        self.consts = torch.zeros(self.cap, 2, dtype=torch.float32, device=dev)
        self._host = torch.zeros(self.cap, 2, dtype=torch.float32).pin_memory()
        self.behind = False                          # some row may be behind step t
        self.flags = None                            # set by the owner: RowFlags of the gradient
        self.pending = None                          # event of the last launch made on another stream
        R.lazy_owners[param.data_ptr()] = self

    def release(self) -> None:
        from . import rasterization as R
        if R.lazy_owners.get(self.param.data_ptr()) is self:
            R.lazy_owners.pop(self.param.data_ptr(), None)

    def record_step(self, step: int, lr: float) -> None:
        """Append the scalars of optimiser step `step` (= self.t + 1) to the device table."""
        if step != self.t + 1:
            raise RuntimeError("LazyRows: steps must be recorded in order")
        self._grow(step)
        _C.check(_C.lib.gags_adam_step_consts(float(lr), self.betas[0], self.betas[1], int(step),
                                              self._host[step].data_ptr()), "gags_adam_step_consts")
        self.consts[step].copy_(self._host[step], non_blocking=True)
        self.t = step

    def _grow(self, step: int) -> None:
        if step < self.cap:
            return
        while step >= self.cap:
            self.cap *= 2
        host = torch.zeros(self.cap, 2, dtype=torch.float32).pin_memory()
        host[:self._host.shape[0]] = self._host
        dev = torch.zeros(self.cap, 2, dtype=torch.float32, device=self.consts.device)
        dev[:self.consts.shape[0]] = self.consts
        # kernels on other streams may still read the old table: keep it alive (a few KB)
        self._retired = getattr(self, "_retired", []) + [self.consts, self._host]
        self._host, self.consts = host, dev

    def record_until(self, step: int, lr: float) -> None:
        """Steps self.t+1 .. step in one go (all rows are current to `step` when this is used: a dense
        pass took those steps, so only the table's length matters, not the entries' lr)."""
        first = self.t + 1
        if step < first:
            return
        self._grow(step)
        for s_ in range(self.t + 1, step + 1):
            _C.check(_C.lib.gags_adam_step_consts(float(lr), self.betas[0], self.betas[1], int(s_),
                                                  self._host[s_].data_ptr()), "gags_adam_step_consts")
        self.consts[first:step + 1].copy_(self._host[first:step + 1], non_blocking=True)
        self.t = step

    def _launch(self, grad, flags, t_to: int, t_apply: int, clear: bool, r0: int = 0,
                r1: int = None) -> None:
        """Rows [r0, r1) only (default: all)."""
        p = self.param
        r1 = self.rows if r1 is None else r1
        if r1 <= r0:
            return
        if self.pending is not None:                 # a step still running on its own stream
            torch.cuda.current_stream(p.device).wait_event(self.pending)
        off = 4 * r0 * self.dim
        _C.check(_C.lib.gags_adam_lazy_rows(
            p.data_ptr() + off, None if grad is None else grad.data_ptr() + off,
            self.exp_avg.data_ptr() + off, self.exp_avg_sq.data_ptr() + off,
            None if flags is None else flags.data_ptr() + r0, self.last.data_ptr() + 4 * r0,
            self.consts.data_ptr(), r1 - r0, self.dim, int(t_to), int(t_apply), self.betas[0],
            self.betas[1], self.eps, 1 if clear else 0, _C.stream_ptr()), "gags_adam_lazy_rows")
        _C.count_launch()

    @torch.no_grad()
    def catch_up(self, flags) -> None:
        """Bring the rows flagged in `flags` (uint8 [N]) up to the current step (before they are read)."""
        if self.behind:
            self._launch(None, flags, self.t, 0, False)

    @torch.no_grad()
    def apply(self, grad, flags, step: int, lr: float, r0: int = 0, r1: int = None) -> None:
        """Optimiser step `step` on the flagged rows only (caught up first, gradient re-zeroed, flags
        cleared); every other row falls one more step behind.  With a row range [r0, r1) the step is
        applied piecewise (SparsePeerAdam overlaps the pieces with the exchange of the next ones):
        the first piece of a step records it."""
        if step == self.t + 1:
            self.record_step(step, lr)
        elif step != self.t:
            raise RuntimeError("LazyRows.apply: steps must be taken in order")
        self._launch(grad, flags, step - 1, step, True, r0, r1)
        self.behind = True

    def wait(self) -> None:
        """Order the current stream behind a step that runs on its own stream."""
        if self.pending is not None:
            torch.cuda.current_stream(self.param.device).wait_event(self.pending)

    @torch.no_grad()
    def flush(self) -> None:
        """Every row up to the current step: afterwards the tensors are what the dense pass holds."""
        self.wait()
        if self.behind:
            self._launch(None, None, self.t, 0, False)
            self.behind = False


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, sparse_rows: bool = False,
                 lazy_rows: bool = False):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.sparse_rows = bool(sparse_rows) or bool(lazy_rows)
        self.lazy_rows = bool(lazy_rows)
        self._rows = {}                     # id(param) -> (persistent grad buffer, RowFlags)
        self._lazy = {}                     # id(param) -> LazyRows
        self._ustreams = {}                 # device index -> stream of the lazily evaluated step
        import os
        self.async_step = os.environ.get("GAGS_B200_ASYNC_ADAM", "1") != "0"

    # ---- row-sparse gradient bookkeeping ---------------------------------------------------------
    def _sparse_ok(self, p) -> bool:
        g = p.grad
        return (self.sparse_rows and p.dim() == 2 and p.shape[1] % 4 == 0 and p.is_contiguous()
                and g is not None and g.is_contiguous() and g.dtype == torch.float32
                and p.dtype == torch.float32 and g.shape == p.shape)

    def _adopt(self, p) -> None:
        """Make the (all-zero) `.grad` buffer of `p` the persistent one: row flags and direct
        reduction by the feature backward.  The backward kernels write through raw pointers and flag
        what they touch; anything that reaches the buffer through torch (autograd's own
        accumulation, a user's in-place op) bumps the tensor's version counter instead, which
        step() takes as "every row may be non-zero"."""
        from . import rasterization as R
        self._release(p)
        rf = R.RowFlags(torch.zeros(p.shape[0], dtype=torch.uint8, device=p.device))
        self._rows[id(p)] = [p.grad, rf, p.grad._version]
        R.row_flags[p.grad.data_ptr()] = rf
        p._gags_direct_grad = True

    def _release(self, p) -> None:
        from . import rasterization as R
        old = self._rows.pop(id(p), None)
        if old is not None:
            R.row_flags.pop(old[0].data_ptr(), None)

    def _check_dense_writes(self, rows) -> None:
        """torch-side writes into the persistent buffer since the last look: flag every row."""
        if rows[0]._version != rows[2]:
            rows[1].flags.fill_(1)
            rows[1].dirty = True
            rows[2] = rows[0]._version

    def __del__(self):
        try:
            from . import rasterization as R
            for rows in self._rows.values():
                R.row_flags.pop(rows[0].data_ptr(), None)
            for lz in self._lazy.values():
                lz.release()
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
                    if (rows[0]._version != rows[2] or rows[1].dirty) and id(p) in self._lazy:
                        self._lazy[id(p)].wait()      # only when device work follows: the usual
                    self._check_dense_writes(rows)    # step(); zero_grad() pair stays asynchronous
                    if rows[1].dirty:
                        p.grad.zero_()
                        rows[1].flags.zero_()
                        rows[1].dirty = False
                        rows[2] = p.grad._version
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
                t, lr = int(st["step"].item()), float(group["lr"])
                lz = self._lazy.get(id(p))
                if lz is not None and lz.param.data_ptr() != p.data_ptr():
                    raise RuntimeError("FusedAdam(lazy_rows): the parameter's storage was replaced; "
                                       "flush() before doing that")
                if rows is not None and rows[0].data_ptr() == p.grad.data_ptr():
                    self._check_dense_writes(rows)
                    if self.lazy_rows:
                        if lz is None:        # every row is current to step t - 1 at this point
                            lz = self._lazy[id(p)] = LazyRows(p, st["exp_avg"], st["exp_avg_sq"],
                                                              (b1, b2), group["eps"])
                            lz.record_until(t - 1, lr)
                            lz.last.fill_(t - 1)
                            lz.flags = rows[1]
                        # The step runs on its own stream: it touches ~7 % of the rows and nothing
                        # the next view's projection / tile sort / weights pass reads, so those
                        # start right behind the backward; the next blend pass and backward wait
                        # for it through rasterization's events, flush() / state_dict() do too.
                        from . import rasterization as R
                        if R.stage_events is not None or not self.async_step:
                            lz.apply(p.grad, rows[1].flags, t, lr)     # (instrumented: in line)
                        else:
                            dev = p.device
                            us = self._ustreams.get(dev.index)
                            if us is None:
                                us = self._ustreams[dev.index] = torch.cuda.Stream(device=dev,
                                                                                   priority=-1)
                            ev0 = torch.cuda.Event()
                            ev0.record(torch.cuda.current_stream(dev))
                            us.wait_event(ev0)
                            with torch.cuda.stream(us):
                                lz.apply(p.grad, rows[1].flags, t, lr)
                                done = torch.cuda.Event()
                                done.record(us)
                            # should the caller drop `.grad` (or this optimiser) right away, the
                            # allocator must not recycle these blocks under the running step
                            for tns in (p.grad, rows[1].flags, lz.last, lz.consts, st["exp_avg"],
                                        st["exp_avg_sq"]):
                                tns.record_stream(us)
                            lz.pending = done
                            R.param_ready_events[p.data_ptr()] = done
                            R.sink_ready_events[p.grad.data_ptr()] = done
                    else:
                        _C.check(_C.lib.gags_adam_step_rows(
                            p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(),
                            st["exp_avg_sq"].data_ptr(), rows[1].flags.data_ptr(), p.shape[0],
                            p.shape[1], lr, float(b1), float(b2), float(group["eps"]), t,
                            _C.stream_ptr()), "gags_adam_step_rows")
                        _C.count_launch()
                    rows[1].dirty = False
                    continue
                if lz is not None:
                    lz.flush()                # a dense step follows: every row must be current
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                zg = zero_grad or sparse
                _C.check(_C.lib.gags_adam_step(p.data_ptr(), g.data_ptr(),
                                               st["exp_avg"].data_ptr(),
                                               st["exp_avg_sq"].data_ptr(), p.numel(), lr, float(b1),
                                               float(b2), float(group["eps"]), t, 1 if zg else 0,
                                               _C.stream_ptr()), "gags_adam_step")
                _C.count_launch()
                if zg and g is not p.grad:
                    p.grad.zero_()
                if lz is not None:            # the dense pass took this step for every row
                    lz.record_until(t, lr)
                    lz.last.fill_(t)
                if sparse:
                    self._adopt(p)            # this (now all-zero) buffer becomes the persistent one
                    if lz is not None:
                        lz.flags = self._rows[id(p)][1]
        return loss

    @torch.no_grad()
    def flush(self) -> None:
        """lazy_rows: bring every row of every parameter up to the current step.  Call before reading
        a parameter or its moments outside render() (checkpoints, export, evaluation code that
        indexes the table directly); state_dict() does it for you."""
        for lz in self._lazy.values():
            lz.flush()

    def state_dict(self):
        self.flush()
        return super().state_dict()

    def load_state_dict(self, state_dict):
        self.flush()
        for lz in self._lazy.values():       # rebuilt against the loaded moments at the next step
            lz.release()
        self._lazy = {}
        return super().load_state_dict(state_dict)

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
        lz = self._lazy.get(id(p))
        if lz is not None:                                 # a dense step: every row current first
            lz.flush()
            lz.record_until(int(st["step"].item()), float(group["lr"]))
            lz.last.fill_(lz.t)
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
