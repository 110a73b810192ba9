#!/usr/bin/env python
"""Exchange-kernel rate probe (run under torchrun on N GPUs): times parallel.PeerAdam's fused
all-reduce + Adam + all-gather kernel alone on a config-3 sized table for both forms (unicast P2P /
NVLS multimem) and several grid sizes, and prints GB/s per direction per rank.
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/peer_rate.py
Drives the tuning hooks gags_set_peer_grid / gags_set_peer_unroll."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GAGS_B200_PEER_TIMING"] = "1"
from gags_b200 import _C, parallel                     # noqa: E402

rank, world, local = parallel.init_from_env()
dev = torch.device("cuda", local)
N, D = 2_000_000, 256
gb = N * D * 4 / 1e9
for form in ("0", "1"):
    os.environ["GAGS_B200_NVLS"] = form
    p = torch.nn.Parameter(torch.zeros(N, D, device=dev))
    peer = parallel.PeerAdam(p, lr=1e-3)
    if form == "1" and not peer.multicast:
        if rank == 0:
            print("no multicast support on this box")
        continue
    # bytes per direction per rank
    per_dir = gb * (1 + 1 / world) if peer.multicast else gb * 2 * (world - 1) / world
    for unroll in ((2, 4, 8) if peer.multicast else (4,)):
        _C.check(_C.lib.gags_set_peer_unroll(unroll))
        for ctas in (1, 2, 4, 8):
            _C.check(_C.lib.gags_set_peer_grid(ctas))
            peer.timing.clear()
            for _ in range(6):
                peer.step()
                peer.synchronize()
            t = peer.timing_summary(last=4)
            if rank == 0:
                print(f"{'NVLS   ' if peer.multicast else 'unicast'} {world} ranks, {ctas} CTAs/SM, "
                      f"{unroll} in flight: kernel {t['kernel_ms']:.3f} ms = "
                      f"{per_dir / t['kernel_ms'] * 1e3:.0f} GB/s per direction "
                      f"(barriers {t['barrier_in_ms']:.3f} + {t['barrier_out_ms']:.3f} ms)", flush=True)
    _C.check(_C.lib.gags_set_peer_grid(0))
    _C.check(_C.lib.gags_set_peer_unroll(4))
    del peer, p
    torch.cuda.synchronize()
    dist.barrier()
dist.destroy_process_group()
