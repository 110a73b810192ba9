"""View-parallel data parallelism.  The reference has no multi-GPU code; the path
shards naturally by training view: one process per GPU, a full replica of the Gaussians on each,
rank r renders views {step*G*k + r*k ... + k-1}, and the only exchange is a sum of
`_semantic_feature.grad` [N,D] over ranks before the (identical) local Adam step."""
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
