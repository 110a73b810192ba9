#!/usr/bin/env python
"""How sparse is the feature gradient of one view?  Config 3, views 0..15: rows of
`_semantic_feature.grad` that the fused loss + backward touched, per view and as a running union
(what a sparse optimiser pass / a sparse multi-GPU exchange could skip).  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gags_b200.arguments import OptimizationParams
from gags_b200.gaussian_renderer import render
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import CONFIGS, config_scene
from gags_b200.utils.loss_utils import l1_backward_fused

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nviews = int(sys.argv[2]) if len(sys.argv) > 2 else 16
dev = torch.device("cuda:0")
n, H, W, D = CONFIGS[cfg]
scene = config_scene(cfg, feature_device=dev)
pc = GaussianModel(3, device=dev)
pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity, scene.features_dc,
                       scene.features_rest, scene.semantic_feature)
pc.training_setup(OptimizationParams(), fused_optimizer=True)
bg = torch.zeros(3, device=dev)
g = torch.Generator().manual_seed(4321)
seg = torch.randint(0, 256, (H // 8 + 1, W // 8 + 1), generator=g, dtype=torch.int32) \
    .repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].contiguous().to(dev)
emb = (0.1 * torch.randn(256, D, generator=g)).to(dev)
union = torch.zeros(n, dtype=torch.bool, device=dev)
for v in range(nviews):
    cam = scene.cameras[v].to(dev)
    pkg = render(cam, pc, None, bg)
    l1_backward_fused(pkg["render"], seg, emb)
    gr = pc._semantic_feature.grad
    touched = gr.abs().amax(dim=1) > 0
    vis = pkg["visibility_filter"]
    union |= touched
    print(f"view {v:2d}: visible {int(vis.sum())/n:.3f}  touched {int(touched.sum())/n:.3f}  "
          f"union(0..{v}) {int(union.sum())/n:.3f}", flush=True)
    pc._semantic_feature.grad = None
