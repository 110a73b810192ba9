import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gags_b200 import _C, rasterization as R
from gags_b200.arguments import OptimizationParams
from gags_b200.gaussian_renderer import render
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import CONFIGS, config_scene, make_target
from gags_b200.utils.loss_utils import l1_loss_fused
dev = torch.device("cuda:0")
n, H, W, D = CONFIGS[3]
scene = config_scene(3)
pc = GaussianModel(3, device=dev)
pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity, scene.features_dc,
                       scene.features_rest, scene.semantic_feature)
pc.training_setup(OptimizationParams(), fused_optimizer=True)
cams = [c.to(dev) for c in scene.cameras]
bg = torch.zeros(3, device=dev)
tgt = make_target(H, W, D, 777).to(dev)
p = pc._semantic_feature
opt = pc.optimizer
from gags_b200.utils.loss_utils import l1_backward_fused
g = torch.Generator().manual_seed(4321)
seg = torch.randint(0, 256, (H // 8 + 1, W // 8 + 1), generator=g, dtype=torch.int32).repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].contiguous().to(dev)
emb = (0.1 * torch.randn(256, D, generator=g)).to(dev)
for step in range(10):
    pkg = render(cams[20 + step], pc, None, bg)
    l1_backward_fused(pkg["render"], seg, emb)
    opt.step(); opt.zero_grad(set_to_none=True)
opt.flush()
for step in range(5):
    pkg = render(cams[40 + step], pc, None, bg)
    l1_backward_fused(pkg["render"], seg, emb)
    pc._semantic_feature.grad = None
torch.cuda.synchronize()
print("main + no-adam legs done; reserved GB", torch.cuda.memory_reserved() / 1e9)
for step in range(10):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    pkg = render(cams[step], pc, None, bg)
    ev[1].record()
    loss = l1_loss_fused(pkg["render"], tgt)
    ev[2].record()
    loss.backward()
    ev[3].record()
    rows = opt._rows.get(id(p))
    info = ""
    if rows is not None:
        info = f"flags {float(rows[1].flags.float().mean()):.3f} ver {rows[0]._version} vs {rows[2]} same_ptr {rows[0].data_ptr() == p.grad.data_ptr()}"
    opt.step()
    opt.zero_grad(set_to_none=True)
    ev[4].record()
    torch.cuda.synchronize()
    t = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    lz = opt._lazy.get(id(p))
    if step == 9:
        t0 = time.time(); opt.flush(); torch.cuda.synchronize(); print("flush ms", (time.time() - t0) * 1e3)
    print(f"step {step}: render {t[0]:.2f} loss {t[1]:.2f} backward {t[2]:.2f} opt {t[3]:.2f} ms | {info} | lazy {lz is not None and lz.behind}")
