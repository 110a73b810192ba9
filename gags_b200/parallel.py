"""View-parallel data parallelism.  The reference has no multi-GPU code; the path
shards naturally by training view: one process per GPU, a full replica of the Gaussians on each,
rank r renders views {step*G*k + r*k ... + k-1}, and the only exchange is a sum of
`_semantic_feature.grad` [N,D] over ranks before the Adam step — as an NCCL all-reduce followed by
the identical local step (allreduce_grads / allreduce_and_step), fused with the step and the
parameter redistribution into one NVLink peer-memory kernel (PeerAdam), or — the default of
bench.py — restricted to the rows the views actually touched (SparsePeerAdam)."""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's env; single-process when WORLD_SIZE is unset."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def views_for_rank(step: int, rank: int, world: int, views_per_rank: int, n_views: int) -> List[int]:
    """Indices of the views rank `rank` renders at optimiser step `step` (disjoint over ranks)."""
    base = step * world * views_per_rank + rank * views_per_rank
    return [(base + j) % n_views for j in range(views_per_rank)]


def allreduce_grads(params: Iterable[torch.Tensor], world: int, async_op: bool = False):
    """Sum .grad over ranks in place (NCCL over NVLink/NVSwitch on GPUs, gloo in CPU tests).
    The sum (not the mean) matches a single process accumulating the same G*k views."""
    handles = []
    if world <= 1:
        return handles
    for p in params:
        if p.grad is None:
            continue
        h = dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, async_op=async_op)
        if async_op:
            handles.append(h)
    return handles


def max_over_ranks(value: float, world: int, device) -> float:
    if world <= 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, world: int, device) -> float:
    if world <= 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def allreduce_and_step(optimizer, param: torch.Tensor, world: int, chunks: int = 8) -> None:
    """Sum `param.grad` over ranks and take the (FusedAdam) optimiser step, pipelined: the gradient
    is all-reduced in `chunks` asynchronous slices and slice i is updated as soon as its reduction
    has landed, while slices i+1.. are still on the wire (NCCL stream).  Equivalent to
    allreduce_grads([param]) followed by optimizer.step()."""
    if world <= 1 or not hasattr(optimizer, "step_chunks") or param.grad is None:
        allreduce_grads([param], world)
        optimizer.step()
        return
    g = param.grad.view(-1)
    numel = g.numel()
    per = -(-numel // chunks)
    per += (-per) % 4
    works = []
    for i in range(chunks):
        s0, s1 = i * per, min(numel, (i + 1) * per)
        if s0 >= s1:
            break
        works.append(dist.all_reduce(g[s0:s1], op=dist.ReduceOp.SUM, async_op=True))

    def ready(i, n):
        if i is None:
            return len(works)
        works[i].wait()            # orders the current stream behind that slice's reduction

    optimizer.step_chunks(param, ready)


def peer_slices(numel: int, world: int) -> tuple[int, int]:
    """(padded element count, elements per rank) of PeerAdam's ownership split: equal contiguous
    slices, each a whole number of float4s, covering `numel`."""
    if numel < 0 or world < 1:
        raise ValueError("numel >= 0 and world >= 1")
    padded = numel + (-numel) % (4 * world)
    return padded, padded // world


class PeerAdam:
    """Gradient all-reduce + Adam + parameter all-gather as ONE kernel over NVLink peer memory
    (`gags_adam_step_peer`, csrc/train_ops.cu) for the view-parallel loop: the baseline is
    allreduce_and_step (NCCL all-reduce, then the full Adam pass on every rank).

    The parameter and its gradient are moved into symmetric memory (every rank maps every rank's
    buffers); rank r owns a contiguous 1/G slice of the table and of the Adam moments.  step():
    barrier (all backward passes done) -> each rank pulls its slice of all G gradients over NVLink,
    sums them in rank order, updates, and pushes the new parameters into all G replicas -> barrier.
    Inbound gradients and outbound parameters use both directions of the links at once, the
    moments' HBM traffic is divided by G, and every replica stays bit-identical.  The gradient
    buffer is persistent: the backward reduces straight into it (rasterization's
    direct_grad_accumulation) and it is re-zeroed on a second stream after the closing barrier.

    Same arithmetic as FusedAdam on the summed gradient (the sum runs in rank order instead of
    NCCL's tree order).  State is sharded: state_dict() holds this rank's slice only."""

    def __init__(self, param: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 group=None):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _C
        from . import rasterization as R
        if not (param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()):
            raise ValueError("PeerAdam needs a contiguous float32 CUDA parameter")
        self._C, self._R = _C, R
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.param = param
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        # per-step CUDA events (wait for the slowest rank / fused kernel / closing barrier): four
        # event records per step on the exchange stream, kept for the last 64 steps
        self.timing = None if os.environ.get("GAGS_B200_PEER_TIMING") == "0" else []
        dev = param.device
        numel = param.numel()
        self.padded, self.per = peer_slices(numel, self.world)
        self.start = self.rank * self.per
        # symmetric buffers: [0] parameters, [1] gradients
        self._buf = symm_mem.empty(2 * self.padded, dtype=torch.float32, device=dev)
        self._hdl = symm_mem.rendezvous(self._buf, self.group)
        self._buf.zero_()
        self._buf[:numel].copy_(param.detach().reshape(-1))
        base = [int(p) for p in self._hdl.buffer_ptrs]
        import ctypes
        self._param_ptrs = (ctypes.c_uint64 * self.world)(*base)
        self._grad_ptrs = (ctypes.c_uint64 * self.world)(*[b + 4 * self.padded for b in base])
        # the module's parameter and its .grad now live in the symmetric buffer
        param.data = self._buf[:numel].view(param.shape)
        self.grad = self._buf[self.padded:self.padded + numel].view(param.shape)
        param.grad = self.grad
        # NVLS: with multicast addresses the switch forms the gradient sum (multimem.ld_reduce) and
        # replicates the parameter stores (multimem.st): per direction a rank then moves 2 GB + 2 GB/G
        # instead of 2 * (G-1)/G * 2 GB — a gain from G > 4 on (GAGS_B200_NVLS=0/1 overrides).
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        want = os.environ.get("GAGS_B200_NVLS", "auto")
        self.multicast = mc != 0 and (want == "1" or (want == "auto" and self.world > 4))
        self._mc_param = mc
        self._mc_grad = mc + 4 * self.padded
        self._xstream = torch.cuda.Stream(device=dev, priority=-1)
        self._done = None
        self.exp_avg = torch.zeros(self.per, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.per, dtype=torch.float32, device=dev)
        R.direct_grad_accumulation = True                 # the backward reduces into `.grad` in place
        torch.cuda.synchronize(dev)
        self._hdl.barrier()

    @torch.no_grad()
    def step(self) -> None:
        """Enqueue the exchange on its own stream behind everything the current stream has queued
        (the backward).  The current stream is free to run the next view's projection / tile sort
        meanwhile — they read no feature — and the next forward blend waits for the new parameters
        (rasterization.param_ready_events); call synchronize() before reading the parameter any
        other way."""
        C, R = self._C, self._R
        p = self.param
        if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr():
            raise RuntimeError("PeerAdam: the parameter's .grad must stay the symmetric buffer "
                               "(use zero_grad(), not set_to_none)")
        self.step_count += 1
        dev = p.device
        main = torch.cuda.current_stream(dev)
        xs = self._xstream
        ev_b = torch.cuda.Event()
        ev_b.record(main)
        xs.wait_event(ev_b)
        with torch.cuda.stream(xs):
            ev = None
            if self.timing is not None:
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                ev[0].record(xs)
            self._hdl.barrier()                           # every rank's backward has finished
            if ev:
                ev[1].record(xs)
            if self.multicast:
                C.check(C.lib.gags_adam_step_multicast(
                    self._mc_grad, self._mc_param, self._buf.data_ptr(), C.ptr(self.exp_avg),
                    C.ptr(self.exp_avg_sq), self.start, self.per, self.lr, self.betas[0],
                    self.betas[1], self.eps, self.step_count, xs.cuda_stream),
                    "gags_adam_step_multicast")
            else:
                C.check(C.lib.gags_adam_step_peer(
                    self.world, self.rank, self._grad_ptrs, self._param_ptrs, C.ptr(self.exp_avg),
                    C.ptr(self.exp_avg_sq), self.start, self.per, self.lr, self.betas[0],
                    self.betas[1], self.eps, self.step_count, xs.cuda_stream),
                    "gags_adam_step_peer")
            C.count_launch()
            if ev:
                ev[2].record(xs)
            self._hdl.barrier()                           # replicas complete, gradients consumed
            if ev:
                ev[3].record(xs)
                self.timing.append(ev)
                if len(self.timing) > 64:
                    del self.timing[:-64]
            # the parameters are final HERE: the next forward blend waits for this event only ...
            params_done = torch.cuda.Event()
            params_done.record(xs)
            # ... while the re-zeroing of the persistent gradient buffer (0.3 ms of HBM writes) only
            # has to land before the next BACKWARD and runs beside the next forward
            C.check(C.lib.gags_zero_fill(self._buf.data_ptr() + 4 * self.padded, 4 * self.padded,
                                         xs.cuda_stream), "gags_zero_fill")
            C.count_launch()
            done = torch.cuda.Event()
            done.record(xs)
        self._done = done
        R.param_ready_events[p.data_ptr()] = params_done  # the next forward blend waits for this
        R.sink_ready_events[self.grad.data_ptr()] = done  # ... and the next backward for this
        p.grad = self.grad

    def synchronize(self) -> None:
        """Order the current stream behind the last exchange (before reading the parameter)."""
        if self._done is not None:
            torch.cuda.current_stream(self.param.device).wait_event(self._done)

    @torch.no_grad()
    def zero_grad(self, set_to_none: bool = False) -> None:
        """step() already re-zeroes the persistent buffer; kept for optimiser-API symmetry."""
        self.param.grad = self.grad

    @torch.no_grad()
    def reset_grad(self) -> None:
        self.synchronize()
        self.grad.zero_()

    def full_moments(self):
        """(exp_avg, exp_avg_sq) of the whole table, gathered from the ranks' shards (flattened)."""
        self.synchronize()
        out = []
        for shard in (self.exp_avg, self.exp_avg_sq):
            parts = [torch.empty_like(shard) for _ in range(self.world)]
            dist.all_gather(parts, shard.contiguous(), group=self.group)
            out.append(torch.cat(parts)[:self.param.numel()].contiguous())
        return out[0], out[1]

    def timing_summary(self, last: int = 10):
        """Average ms of (wait for the slowest rank, fused kernel, closing barrier) over the last
        steps (GAGS_B200_PEER_TIMING=0 switches the event records off)."""
        if not self.timing:
            return None
        torch.cuda.synchronize(self.param.device)
        ev = self.timing[-last:]
        n = len(ev)
        return {"barrier_in_ms": sum(e[0].elapsed_time(e[1]) for e in ev) / n,
                "kernel_ms": sum(e[1].elapsed_time(e[2]) for e in ev) / n,
                "barrier_out_ms": sum(e[2].elapsed_time(e[3]) for e in ev) / n}

    def state_dict(self):
        return {"step": self.step_count, "rank": self.rank, "world": self.world,
                "exp_avg_shard": self.exp_avg, "exp_avg_sq_shard": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps}


def row_word_ranges(rows: int, world: int):
    """Ownership split of the row-sparse exchange (csrc/train_ops.cu gags_grad_allreduce_rows):
    rows are handled in words of 4 (one 32-bit word of uint8 flags); rank r owns words
    [r * per, (r + 1) * per) clipped to the word count.  Returns (words, [(w0, w1) per rank])."""
    if rows < 0 or world < 1:
        raise ValueError("rows >= 0 and world >= 1")
    words = (rows + 3) // 4
    per = -(-words // world) if words else 0
    return words, [(min(words, r * per), min(words, (r + 1) * per)) for r in range(world)]


class SparsePeerAdam:
    """View-parallel optimiser step that exchanges only what the views touched.

    One view's feature backward reaches a small part of the [N, D] table (7 % of the rows at
    BASELINE config 3; 15 % for the union over 8 consecutive views), and the backward already flags
    those rows (rasterization.row_flags).  The gradient buffer and the flags live in symmetric
    memory.  step(): barrier -> ONE kernel sums, over NVLink peer memory (NVLS multimem where the
    fabric offers it), only the rows flagged on some rank and writes the sums into every replica
    in place (gags_grad_allreduce_rows) -> barrier -> every rank's own row-sparse Adam pass
    (gags_adam_step_rows: flagged rows read + re-zeroed, every other row takes the g = 0 update).
    Each rank keeps the FULL optimiser state, so no parameter travels at all: per step and rank the
    fabric moves ~0.35 GB per direction instead of the dense exchange's 2 x 2.25 GB (PeerAdam),
    which stays available for gradients that are not row sparse.  With `lazy` (default) the local
    pass is the lazily evaluated one (optim.LazyRows): it visits the union rows only, and rows no
    rank touched take their zero-gradient steps when they are next needed.  Same arithmetic as FusedAdam on
    the summed gradient; replicas stay bit-identical because exactly one rank forms each row's sum
    and all ranks apply the same elementwise update to it."""

    def __init__(self, param: torch.Tensor, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 group=None, lazy=None):
        import ctypes
        import torch.distributed._symmetric_memory as symm_mem
        from . import _C
        from . import rasterization as R
        if lazy is None:
            lazy = os.environ.get("GAGS_B200_LAZY_ADAM", "1") != "0"
        if not (param.is_cuda and param.dtype == torch.float32 and param.is_contiguous()
                and param.dim() == 2 and param.shape[1] % 4 == 0):
            raise ValueError("SparsePeerAdam needs a contiguous float32 CUDA [N, D] parameter, D % 4 == 0")
        self._C, self._R = _C, R
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.param = param
        self.lr, self.betas, self.eps = float(lr), (float(betas[0]), float(betas[1])), float(eps)
        self.step_count = 0
        self.timing = None if os.environ.get("GAGS_B200_PEER_TIMING") == "0" else []
        dev = param.device
        self.rows, self.dim = int(param.shape[0]), int(param.shape[1])
        numel = param.numel()
        self.words, _ = row_word_ranges(self.rows, self.world)
        # symmetric buffer: the gradient [N, D], then one 32-bit word of row flags per 4 rows
        self._buf = symm_mem.empty(numel + self.words, dtype=torch.float32, device=dev)
        self._hdl = symm_mem.rendezvous(self._buf, self.group)
        self._buf.zero_()
        base = [int(p) for p in self._hdl.buffer_ptrs]
        self._grad_ptrs = (ctypes.c_uint64 * self.world)(*base)
        self._flag_ptrs = (ctypes.c_uint64 * self.world)(*[b + 4 * numel for b in base])
        self.grad = self._buf[:numel].view(param.shape)
        self._flag_bytes = self._buf[numel:].view(torch.uint8)           # 4 * words bytes
        self.flags = self._flag_bytes[:self.rows]
        self.union_flags = torch.zeros(4 * self.words, dtype=torch.uint8, device=dev)
        mc = int(getattr(self._hdl, "multicast_ptr", 0) or 0)
        want = os.environ.get("GAGS_B200_NVLS", "auto")
        self.multicast = mc != 0 and (want == "1" or (want == "auto" and self.world > 2))
        self._mc_grad = mc if self.multicast else None
        self._mc_flags = (mc + 4 * numel) if self.multicast else None
        self._xstream = torch.cuda.Stream(device=dev, priority=-1)
        self._ustream = torch.cuda.Stream(device=dev, priority=-1)       # the local update pass
        self._done = None
        # row ranges processed as a pipeline (exchange of range c+1 beside the update of range c);
        # boundaries on flag words, per-range pointer tables for the peers' buffers
        self.chunks = max(1, int(os.environ.get("GAGS_B200_EXCHANGE_CHUNKS", "4")))
        wc = -(-self.words // self.chunks) if self.words else 0
        self._chunk_args = []
        for c in range(self.chunks):
            r0, r1 = min(self.rows, 4 * wc * c), min(self.rows, 4 * wc * (c + 1))
            if r1 <= r0:
                continue
            gp = (ctypes.c_uint64 * self.world)(*[b + 4 * r0 * self.dim for b in base])
            fp = (ctypes.c_uint64 * self.world)(*[b + 4 * numel + r0 for b in base])
            self._chunk_args.append((r0, r1, gp, fp))
        self.exp_avg = torch.zeros_like(param, memory_format=torch.preserve_format)
        self.exp_avg_sq = torch.zeros_like(param, memory_format=torch.preserve_format)
        param.grad = self.grad
        self._rf = R.RowFlags(self.flags)
        R.row_flags[self.grad.data_ptr()] = self._rf       # the backward flags the rows it touches
        param._gags_direct_grad = True                     # ... and reduces into `.grad` in place
        # lazily evaluated update (optim.LazyRows): the local pass visits the union rows only, and
        # the two-pass forward brings the rows a view reads up to date before reading them
        self.lz = None
        if lazy:
            from .optim import LazyRows
            self.lz = LazyRows(param, self.exp_avg, self.exp_avg_sq, self.betas, self.eps)
            self.lz.flags = self._rf
        torch.cuda.synchronize(dev)
        self._hdl.barrier()

    @torch.no_grad()
    def step(self) -> None:
        """Enqueue exchange + update on their own streams behind everything the current stream has
        queued (the backward); the current stream goes on with the next view's projection / tile
        sort / weights pass, and the next blend pass / backward wait for the update through
        rasterization.param_ready_events / sink_ready_events.  The table is processed in
        `self.chunks` row ranges: while the fabric sums the rows of range c+1, the HBM-bound local
        Adam pass already runs on range c (a second stream)."""
        C, R = self._C, self._R
        p = self.param
        if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr():
            raise RuntimeError("SparsePeerAdam: the parameter's .grad must stay the symmetric buffer "
                               "(use zero_grad(), not set_to_none)")
        self.step_count += 1
        dev = p.device
        main = torch.cuda.current_stream(dev)
        xs, us = self._xstream, self._ustream
        ev_b = torch.cuda.Event()
        ev_b.record(main)
        xs.wait_event(ev_b)
        ev = None
        if self.timing is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        with torch.cuda.stream(xs):
            if ev:
                ev[0].record(xs)
            self._hdl.barrier()                           # every rank's backward has finished
            if ev:
                ev[1].record(xs)
        for c, (r0, r1, gp, fp) in enumerate(self._chunk_args):
            last = c == len(self._chunk_args) - 1
            with torch.cuda.stream(xs):
                off = 4 * r0 * self.dim
                C.check(C.lib.gags_grad_allreduce_rows(
                    self.world, self.rank, gp, fp,
                    None if self._mc_grad is None else self._mc_grad + off,
                    None if self._mc_flags is None else self._mc_flags + r0,
                    self.union_flags.data_ptr() + r0, r1 - r0, self.dim, xs.cuda_stream),
                    "gags_grad_allreduce_rows")
                C.count_launch()
                if ev and last:
                    ev[2].record(xs)
                self._hdl.barrier()                       # every replica holds this range's sums
                if ev and last:
                    ev[3].record(xs)
                summed = torch.cuda.Event()
                summed.record(xs)
                if last:                                  # every rank has read every rank's flags
                    C.check(C.lib.gags_memset_zero(C.ptr(self._flag_bytes), 4 * self.words,
                                                   xs.cuda_stream), "gags_memset_zero")
                    cleared = torch.cuda.Event()
                    cleared.record(xs)
            us.wait_event(summed)
            with torch.cuda.stream(us):
                if self.lz is not None:
                    self.lz.apply(self.grad, self.union_flags, self.step_count, self.lr, r0, r1)
                else:
                    off = 4 * r0 * self.dim
                    C.check(C.lib.gags_adam_step_rows(
                        p.data_ptr() + off, self.grad.data_ptr() + off,
                        self.exp_avg.data_ptr() + off, self.exp_avg_sq.data_ptr() + off,
                        self.union_flags.data_ptr() + r0, r1 - r0, self.dim, self.lr, self.betas[0],
                        self.betas[1], self.eps, self.step_count, us.cuda_stream),
                        "gags_adam_step_rows")
                    C.count_launch()
        us.wait_event(cleared)
        with torch.cuda.stream(us):
            if ev:
                ev[4].record(us)
                self.timing.append(ev)
                if len(self.timing) > 64:
                    del self.timing[:-64]
            done = torch.cuda.Event()
            done.record(us)
        self._done = done
        self._rf.dirty = False
        R.param_ready_events[p.data_ptr()] = done         # the next blend pass waits for this
        R.sink_ready_events[self.grad.data_ptr()] = done  # ... and so does the next backward
        p.grad = self.grad

    def synchronize(self) -> None:
        if self._done is not None:
            torch.cuda.current_stream(self.param.device).wait_event(self._done)

    @torch.no_grad()
    def zero_grad(self, set_to_none: bool = False) -> None:
        """step() re-zeroes the rows it consumed; a gradient that was never applied is dropped here."""
        self.param.grad = self.grad
        if self._rf.dirty:
            self.reset_grad()

    @torch.no_grad()
    def reset_grad(self) -> None:
        self.synchronize()
        self.grad.zero_()
        self._flag_bytes.zero_()
        self._rf.dirty = False

    @torch.no_grad()
    def flush(self) -> None:
        """Lazy mode: bring every row up to the current step (before reading the parameter or the
        moments outside render())."""
        self.synchronize()
        if self.lz is not None:
            self.lz.flush()

    def full_moments(self):
        """(exp_avg, exp_avg_sq) of the whole table, flattened copies (every rank holds them)."""
        self.flush()
        return self.exp_avg.detach().clone().view(-1), self.exp_avg_sq.detach().clone().view(-1)

    def timing_summary(self, last: int = 10):
        if not self.timing:
            return None
        torch.cuda.synchronize(self.param.device)
        ev = self.timing[-last:]
        n = len(ev)
        return {"barrier_in_ms": sum(e[0].elapsed_time(e[1]) for e in ev) / n,
                "kernel_ms": sum(e[1].elapsed_time(e[2]) for e in ev) / n,
                "barrier_out_ms": sum(e[2].elapsed_time(e[3]) for e in ev) / n,
                "adam_ms": sum(e[3].elapsed_time(e[4]) for e in ev) / n}

    def state_dict(self):
        self.flush()
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "lr": self.lr, "betas": self.betas, "eps": self.eps}
