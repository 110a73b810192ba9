#!/usr/bin/env python
"""Kernel-level timeline of a few training steps (torch.profiler / CUPTI): start, duration, stream of
every kernel, to see what overlaps with what.  Run on the GPU box; prints a compact table."""
import json, os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import profile, ProfilerActivity
from gags_b200 import rasterization as R
from gags_b200.arguments import OptimizationParams
from gags_b200.gaussian_renderer import render
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import config_scene
from gags_b200.utils.loss_utils import l1_loss_segmap_fused, l1_backward_fused
dev = torch.device("cuda:0")
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
def step(i):
    pkg = render(cams[i % 64], pc, None, bg)
    if os.environ.get("UNFUSED_LOSS"):
        loss = l1_loss_segmap_fused(pkg["render"], seg, emb)
        loss.backward()
    else:
        l1_backward_fused(pkg["render"], seg, emb)
    pc.optimizer.step()
    pc.optimizer.zero_grad(set_to_none=True)
for i in range(10): step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(10, 14): step(i)
    torch.cuda.synchronize()
path = os.path.join(tempfile.gettempdir(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0 = ev[0]["ts"]
for e in ev:
    name = e["name"].replace("void ", "").replace("(anonymous namespace)::", "")[:44]
    print(f"{(e['ts']-t0)/1000:9.3f} ms  +{e['dur']/1000:7.3f}  s{e['args'].get('stream')}  {name}")
