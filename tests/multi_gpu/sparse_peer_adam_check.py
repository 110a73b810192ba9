"""Run under torchrun on >= 2 GPUs: parallel.SparsePeerAdam (row-sparse gradient all-reduce over
peer memory + row-sparse Adam on every rank) against dist.all_reduce + FusedAdam on the same random
row-sparse gradients, three steps, two table shapes; parameters and moments must match, replicas
must end bit-identical, the gradient buffer and the flags must come back all zero."""
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
ok_all = True
for N, D, frac, lazy in ((10007, 64, 0.1, False), (10007, 64, 0.1, True), (4099, 12, 0.5, True),
                         (3, 256, 1.0, True)):
    torch.manual_seed(0)
    p0 = torch.randn(N, D, device=dev)
    ref = torch.nn.Parameter(p0.clone())
    opt = FusedAdam([ref], lr=1e-2)
    par = torch.nn.Parameter(p0.clone())
    peer = parallel.SparsePeerAdam(par, lr=1e-2, lazy=lazy)
    for step in range(4):
        g = torch.Generator(device=dev).manual_seed(100 * step + rank + 7 * N)
        touched = torch.rand(N, device=dev, generator=g) < frac
        grad = torch.randn(N, D, device=dev, generator=g) * touched[:, None].float()
        ref.grad = grad.clone()
        dist.all_reduce(ref.grad, op=dist.ReduceOp.SUM)
        opt.step()
        # what the feature backward does: wait for the buffer, reduce into it in place, flag rows
        ev = R.sink_ready_events.pop(par.grad.data_ptr(), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)
        par.grad.add_(grad)
        peer.flags.copy_(touched.to(torch.uint8))
        peer.step()
        peer.synchronize()
    behind = lazy and frac < 1.0 and not torch.equal(par.detach(), ref.detach())
    peer.flush()                                     # lazy: every row up to the current step
    torch.cuda.synchronize()
    scale = ref.detach().abs().max()
    err = float((par.detach() - ref.detach()).abs().max() / scale)
    st = opt.state[ref]
    err_m = float((peer.exp_avg - st["exp_avg"]).abs().max() / st["exp_avg"].abs().max())
    err_v = float((peer.exp_avg_sq - st["exp_avg_sq"]).abs().max() / st["exp_avg_sq"].abs().max())
    clean = float(par.grad.abs().max()) == 0.0 and int(peer._flag_bytes.sum()) == 0 \
        and int(peer.union_flags.sum()) == 0
    chk = par.detach().view(torch.int32).to(torch.int64).sum().reshape(1)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(int(c) == int(allc[0]) for c in allc)
    ok = err < 1e-6 and err_m < 1e-6 and err_v < 1e-6 and same and clean \
        and (behind or not lazy or frac >= 1.0)
    ok_all = ok_all and ok
    print(f"rank {rank}: [{N},{D}] lazy={lazy} multicast={peer.multicast} rel err p {err:.2e} m {err_m:.2e} v {err_v:.2e}, "
          f"replicas identical {same}, buffers clean {clean} -> {'OK' if ok else 'FAIL'}", flush=True)
    if peer.lz is not None:
        peer.lz.release()
    del peer
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok_all else 1)
