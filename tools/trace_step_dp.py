#!/usr/bin/env python
"""Kernel-level timeline of the view-parallel training step (torch.profiler / CUPTI) on rank 0:
start, duration, stream of every kernel of four steps with parallel.SparsePeerAdam (TRACE_DENSE=1: PeerAdam).
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/trace_step_dp.py"""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
from torch.profiler import profile, ProfilerActivity
from gags_b200 import parallel, rasterization as R
from gags_b200.arguments import OptimizationParams
from gags_b200.gaussian_renderer import render
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import config_scene
from gags_b200.utils.loss_utils import l1_backward_fused
rank, world, local = parallel.init_from_env("nccl")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
scene = config_scene(3)
pc = GaussianModel(3, device=dev)
pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity, scene.features_dc,
                       scene.features_rest, scene.semantic_feature)
pc.training_setup(OptimizationParams(), fused_optimizer=True)
cams = [c.to(dev) for c in scene.cameras]
R.register_static(*[c.world_view_transform for c in cams])
bg = torch.zeros(3, device=dev)
g = torch.Generator().manual_seed(1)
seg = torch.randint(0, 256, (1080, 1920), generator=g, dtype=torch.int32).to(dev)
emb = (0.1 * torch.randn(256, 256, generator=g)).to(dev)
grp = pc.optimizer.param_groups[0]
peer = (parallel.PeerAdam if os.environ.get("TRACE_DENSE") else parallel.SparsePeerAdam)(pc._semantic_feature, lr=grp["lr"], betas=grp["betas"], eps=grp["eps"])
def step(i):
    v = parallel.views_for_rank(i, rank, world, 1, 64)[0]
    pkg = render(cams[v], pc, None, bg)
    l1_backward_fused(pkg["render"], seg, emb)
    peer.step()
for i in range(10): step(i)
peer.synchronize(); torch.cuda.synchronize(); dist.barrier()
acts = [ProfilerActivity.CUDA, ProfilerActivity.CPU]
# every rank profiles itself; rank 0 prints, all ranks also write gpurun_out/trace_dp{world}_rank{r}.txt
with profile(activities=acts) as prof:
    for i in range(10, 14): step(i)
    peer.synchronize(); torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), f"trace_dp_{rank}.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
lines = []
for e in ev:
    name = e["name"].replace("void ", "").replace("(anonymous namespace)::", "")[:48]
    lines.append(f"{(e['ts']-t0)/1000:9.3f} ms  +{e['dur']/1000:7.3f}  s{e['args'].get('stream')}  {name}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", f"trace_dp{world}_rank{rank}.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
if rank == 0:
    print("\n".join(lines))
dist.barrier()
dist.destroy_process_group()
