"""Run under torchrun on >= 2 GPUs (tools/gpu_multi.sh): parallel.PeerAdam (all-reduce + Adam +
all-gather in one peer-memory kernel) against dist.all_reduce + FusedAdam on the same random
gradients, three steps; replicas must end bit-identical across ranks."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from gags_b200 import parallel                      # noqa: E402
from gags_b200 import rasterization as R            # noqa: E402
from gags_b200.optim import FusedAdam               # noqa: E402

rank, world, local = parallel.init_from_env()
dev = torch.device("cuda", local)
torch.manual_seed(0)
N, D = 10007, 64                                    # numel not a multiple of 4 * world
p0 = torch.randn(N, D, device=dev)
ref = torch.nn.Parameter(p0.clone())
opt = FusedAdam([ref], lr=1e-2)
par = torch.nn.Parameter(p0.clone())
peer = parallel.PeerAdam(par, lr=1e-2)
for step in range(3):
    g = torch.Generator(device=dev).manual_seed(100 * step + rank)
    grad = torch.randn(N, D, device=dev, generator=g)
    grad[::3] = 0.0
    # baseline: NCCL sum + full local Adam
    ref.grad = grad.clone()
    dist.all_reduce(ref.grad, op=dist.ReduceOp.SUM)
    opt.step()
    # fused: the backward would reduce into par.grad in place
    ev = R.sink_ready_events.pop(par.grad.data_ptr(), None)
    if ev is not None:                               # the re-zeroing of the persistent buffer
        torch.cuda.current_stream().wait_event(ev)
    par.grad.add_(grad)
    peer.step()
    peer.synchronize()
torch.cuda.synchronize()
err = float((par.detach() - ref.detach()).abs().max() / ref.detach().abs().max())
# replicas identical across ranks
chk = par.detach().double().sum().reshape(1)
allc = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(allc, chk)
same = all(float(c) == float(allc[0]) for c in allc)
ok = err < 1e-6 and same
print(f"rank {rank}: multicast={peer.multicast} rel err vs all_reduce + FusedAdam {err:.3e}, replicas identical {same} -> "
      f"{'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
