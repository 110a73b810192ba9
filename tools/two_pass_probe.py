#!/usr/bin/env python
"""Two-pass forward (weights pass + blend-from-cache pass) against the single-pass cached forward on a
BASELINE config: identical bits expected; CUDA-event times of each.  Run on the GPU box."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gags_b200 import _C, rasterization as R
from gags_b200.synthetic import CONFIGS, config_scene

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
n, H, W, D = CONFIGS[cfg]
scene = config_scene(cfg, feature_device=dev)
cam0 = scene.cameras[0].to(dev)
fx = W / (2 * math.tan(cam0.FoVx * 0.5)); fy = H / (2 * math.tan(cam0.FoVy * 0.5))
tw, th = (W + 15) // 16, (H + 15) // 16
flags = _C.GAGS_F_LOG_SCALES | _C.GAGS_F_LOGIT_OPACITY
cam, keep = R.make_camera(cam0.world_view_transform.T.contiguous(), fx, fy, W / 2, H / 2, W, H, flags=flags)
g = lambda t: t.to(dev).contiguous()
radii, m2d, dep, con, opac, tiles, geom = R._Project.apply(g(scene.xyz), g(scene.rotation), g(scene.scaling),
                                                            g(scene.opacity).reshape(-1), cam, keep, tw, th)
b = R.bin_and_sort(m2d, radii, dep, tiles, tw, th)
offsets, ids = b["offsets"], b["flatten_ids"]
cols = g(scene.semantic_feature)
n_half = tw * ((H + 7) // 8)
slots = int(_C.lib.gags_blend_cache_slots(ids.numel(), tw * th))
def bufs():
    return (torch.empty(slots * 16384, dtype=torch.uint8, device=dev), torch.empty(slots * 32, dtype=torch.int32, device=dev),
            torch.empty(slots, dtype=torch.int32, device=dev), torch.zeros(n_half + 1, dtype=torch.int32, device=dev))
c0, c1 = bufs(), bufs()
st = _C.stream_ptr()
def single():
    r = torch.empty(H, W, D, device=dev); a = torch.empty(H, W, device=dev)
    _C.check(_C.lib.gags_blend_fwd_cached(_C.ptr(geom), _C.ptr(cols), D, None, W, H, _C.ptr(offsets), _C.ptr(ids),
                                          _C.ptr(r), _C.ptr(a), None, *[_C.ptr(x) for x in c0], st))
    return r, a
def weights():
    a = torch.empty(H, W, device=dev)
    _C.check(_C.lib.gags_blend_fwd_weights(_C.ptr(geom), W, H, _C.ptr(offsets), _C.ptr(ids), _C.ptr(a), None,
                                           *[_C.ptr(x) for x in c1], st))
    return a
def blend(a):
    r = torch.empty(H, W, D, device=dev)
    _C.check(_C.lib.gags_blend_fwd_from_cache(_C.ptr(cols), D, None, W, H, _C.ptr(offsets), *[_C.ptr(x) for x in c1],
                                              _C.ptr(a), _C.ptr(r), st))
    return r
r0, a0 = single()
a1 = weights()
r1 = blend(a1)
torch.cuda.synchronize()
print("alphas equal", torch.equal(a0, a1), "render equal", torch.equal(r0, r1),
      "max diff", float((r0 - r1).abs().max()), "counts equal", torch.equal(c0[3][:n_half], c1[3][:n_half]),
      "batches", int(c1[3][:n_half].sum()))
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]
print("single-pass cached forward  min/median ms", timeit(lambda: single()))
print("weights pass                min/median ms", timeit(lambda: weights()))
print("blend-from-cache pass       min/median ms", timeit(lambda: blend(a1)))
