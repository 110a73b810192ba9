#!/usr/bin/env python
"""Per-role cycle timeline of a few CTAs of blend_fwd_tc on the config-3 scene (debug library built
with `python -m gags_b200.build --timing`).  Run on the GPU box."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["GAGS_B200_LIB"] = os.path.join(ROOT, "gags_b200", "csrc", "libgags_b200_dbg.so")
import torch
from gags_b200 import _C
from gags_b200.gaussian_renderer import render
from gags_b200.scene import GaussianModel
from gags_b200.synthetic import config_scene

dev = torch.device("cuda:0")
scene = config_scene(3)
pc = GaussianModel(3, device=dev)
pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity, scene.features_dc,
                       scene.features_rest, scene.semantic_feature)
cam = scene.cameras[0].to(dev)
bg = torch.zeros(3, device=dev)
with torch.no_grad():
    for _ in range(3):
        render(cam, pc, None, bg)
torch.cuda.synchronize()
# one training-mode pass for the cached backward timeline
from gags_b200.arguments import OptimizationParams
pc.training_setup(OptimizationParams(), fused_optimizer=True)      # freezes the geometry
pkg = render(cam, pc, None, bg)
pkg["render"].sum().backward()
torch.cuda.synchronize()
nb = 8 * 4 * 16 * 8
bufb = (ctypes.c_longlong * nb)()
_C.lib.gags_debug_timeline_bwd.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert _C.lib.gags_debug_timeline_bwd(bufb, nb) == 0
tbw = torch.tensor(list(bufb), dtype=torch.int64).reshape(8, 4, 16, 8)
print("BACKWARD (cached) columns: load(issue) | mma(top wfull_ok accfree_ok committed) | epi(top accfull_ok drained reduced)")
for slot in range(8):
    t0 = int(tbw[slot, 3, 0, 0])
    if t0 == 0:
        continue
    print(f"=== bwd CTA slot {slot} (persistent; steps run across jobs): vfull job0 +{int(tbw[slot,3,0,1])-t0}, "
          f"vfull job1 +{int(tbw[slot,3,0,2])-t0}, CTA end +{int(tbw[slot,3,0,3])-t0}")
    for b in range(16):
        if tbw[slot, 2, b, 0] == 0:
            break
        f = lambda role, k: (int(tbw[slot, role, b, k]) - t0) if tbw[slot, role, b, k] else -1
        print(f" s{b:2d} | load {f(1,0):6d} | mma " + " ".join(f"{f(2,k):6d}" for k in range(4)) +
              " | epi " + " ".join(f"{f(0,k):6d}" for k in range(4)))

n = 8 * 4 * 64 * 8
buf = (ctypes.c_longlong * n)()
_C.lib.gags_debug_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
rc = _C.lib.gags_debug_timeline(buf, n)
assert rc == 0, rc
t = torch.tensor(list(buf), dtype=torch.int64).reshape(8, 4, 64, 8)
names = {0: ["top", "list_ok", "computed", "stored"], 1: ["top", None, "free_ok", "half", "stored"],
         2: ["wait_full", "full_ok", "committed"], 3: ["top", "scanned", "published"]}
print("columns: scan(top scanned published) | conv(top free_ok half stored) | pix(top list_ok computed stored) | mma(wait_full full_ok committed)")
for slot in range(8):
    t0 = int(t[slot, 3, 0, 0])
    if t0 == 0:
        continue
    print(f"=== CTA slot {slot}: epilogue-sync at +{int(t[slot,3,0,1])-t0}, after-sync +{int(t[slot,3,0,2])-t0}, end +{int(t[slot,3,0,3])-t0}")
    for b in range(62):
        if t[slot, 0, b, 0] == 0 and t[slot, 3, b + 1, 0] == 0:
            break
        line = f" b{b:2d} "
        for role, tag, bb in ((3, "scan", b + 1), (1, "conv", b), (0, "pix", b), (2, "mma", b)):
            vals = [int(t[slot, role, bb, k]) - t0 if t[slot, role, bb, k] else -1
                    for k, nm in enumerate(names[role]) if nm is not None]
            line += f"| {tag} " + " ".join(f"{v:6d}" for v in vals) + " "
        print(line)
