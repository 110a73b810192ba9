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
from gags_b200.utils.loss_utils import l1_backward_fused
dev = torch.device("cuda:0")
D, H, W = 64, 72, 112
scene = make_scene(6000, H, W, D, seed=33, n_views=8, sigma_px_median=1.5)
g = torch.Generator().manual_seed(6)
seg = torch.randint(0, 9, (H, W), generator=g, dtype=torch.int32).to(dev)
emb = (0.2 * torch.randn(9, D, generator=g)).to(dev)
bg = torch.zeros(3, device=dev)
def model():
    pc = GaussianModel(3, device=dev)
    pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity,
                           scene.features_dc, scene.features_rest, scene.semantic_feature)
    pc.training_setup(OptimizationParams(), fused_optimizer=True)
    return pc
pc, shadow = model(), model()
p = pc._semantic_feature
opt = pc.optimizer = FusedAdam([{"params": [p], "lr": 1e-2}], lr=1e-2, eps=1e-15, lazy_rows=True)
ps = shadow._semantic_feature
ms, vs = torch.zeros_like(ps), torch.zeros_like(ps)
ever = torch.zeros(6000, dtype=torch.bool, device=dev)
for it in range(1, 7):
    cam = scene.cameras[(it - 1) % 8].to(dev)
    pkg = render(cam, pc, None, bg)
    with torch.no_grad():
        ref = render(cam, shadow, None, bg)["render"]
    print(it, "render equal", torch.equal(pkg["render"].detach(), ref))
    l1_backward_fused(pkg["render"], seg, emb)
    gcopy = p.grad.clone()
    rows = opt._rows.get(id(p))
    if rows is not None:
        fl = rows[1].flags.bool()
        nz = gcopy.abs().amax(1) > 0
        print("   flagged", int(fl.sum()), "nonzero rows", int(nz.sum()), "nz not flagged", int((nz & ~fl).sum()))
        ever |= fl
    opt.step(); opt.zero_grad(set_to_none=True)
    _C.check(_C.lib.gags_adam_step(ps.data_ptr(), gcopy.data_ptr(), ms.data_ptr(), vs.data_ptr(),
                                   ps.numel(), 1e-2, 0.9, 0.999, 1e-15, it, 0, _C.stream_ptr()))
    torch.cuda.synchronize()
    lz = opt._lazy.get(id(p))
    if lz is not None:
        print("   lazy t", lz.t, "behind", lz.behind, "last hist", torch.bincount(lz.last, minlength=it + 1).tolist(),
              "consts", lz.consts[:it + 1].tolist())
lz = opt._lazy[id(p)]
print("before eval: behind", lz.behind, "owners", list(R.lazy_owners.keys()), "p ptr", p.data_ptr())
with torch.no_grad():
    cam = scene.cameras[5].to(dev)
    a = render(cam, pc, None, bg)["render"]
    b = render(cam, shadow, None, bg)["render"]
print("eval equal", torch.equal(a, b), "behind after", lz.behind, "last hist", torch.bincount(lz.last, minlength=7).tolist())
torch.cuda.synchronize()
st = opt.state[p]
bad = (p.detach() != ps.detach()).any(1)
print("rows differing after flush", int(bad.sum()), "of", p.shape[0], "; of those ever flagged", int((bad & ever).sum()))
badm = (st["exp_avg"] != ms).any(1); badv = (st["exp_avg_sq"] != vs).any(1)
print("m rows differing", int(badm.sum()), "v rows differing", int(badv.sum()))
if bad.any():
    i = int(bad.nonzero()[0])
    print("row", i, "last", int(lz.last[i]), "ever", bool(ever[i]))
    print(" p lazy", p.detach()[i, :4].tolist(), " dense", ps.detach()[i, :4].tolist())
    print(" m lazy", st["exp_avg"][i, :4].tolist(), " dense", ms[i, :4].tolist())
    print(" v lazy", st["exp_avg_sq"][i, :4].tolist(), " dense", vs[i, :4].tolist())
    print(" max abs diff p", float((p.detach() - ps.detach()).abs().max()))
