import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gags_b200 import _C, rasterization as R
from gags_b200.arguments import OptimizationParams
from gags_b200.gaussian_renderer import render
from gags_b200.optim import FusedAdam
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import make_scene
from gags_b200.utils.loss_utils import l1_loss_segmap_fused
dev = torch.device("cuda:0")
D, H, W = 128, 72, 112
scene = make_scene(6000, H, W, D, seed=21, n_views=8, sigma_px_median=1.5)
g = torch.Generator().manual_seed(5)
seg = torch.randint(0, 9, (H, W), generator=g, dtype=torch.int32).to(dev)
emb = (0.2 * torch.randn(9, D, generator=g)).to(dev)
bg = torch.zeros(3, device=dev)
pc = GaussianModel(3, device=dev)
pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity,
                       scene.features_dc, scene.features_rest, scene.semantic_feature)
pc.training_setup(OptimizationParams(), fused_optimizer=True)
p = pc._semantic_feature
opt = pc.optimizer = FusedAdam([{"params": [p], "lr": 1e-2}], lr=1e-2, eps=1e-15, sparse_rows=True)
orig_hook = opt._on_autograd_accumulate
def hook(q):
    print("   HOOK fired; grad ptr", q.grad.data_ptr() if q.grad is not None else None)
    return orig_hook(q)
opt._on_autograd_accumulate = hook
orig_mark = R._mark_rows
def mark(v, cache, offsets, w, h):
    print("   _mark_rows: cache is None?", cache is None, "registered", v is not None and v.data_ptr() in R.row_flags)
    return orig_mark(v, cache, offsets, w, h)
R._mark_rows = mark
for it in range(3):
    pkg = render(scene.cameras[it].to(dev), pc, None, bg)
    print("it", it, "fused handle", getattr(pkg["render"], "_gags_fused", None) is not None)
    l1_loss_segmap_fused(pkg["render"], seg, emb).backward()
    rows = opt._rows.get(id(p))
    if rows is not None:
        print("   flags frac", float(rows[1].flags.float().mean()), "nz rows", float((p.grad.abs().amax(1) > 0).float().mean()))
    opt.step(); opt.zero_grad(set_to_none=True)
