#!/usr/bin/env python
"""CPU-side statistics of a benchmark workload (uses the oracle; dev tool, not product code):
per 16x8 half tile - list length, survivors of the alpha >= 1/255 bounding-box cull, where the
half tile terminates, contributors per pixel.  python tools/workload_stats.py [config] [n_tiles]"""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gags_b200.synthetic import CONFIGS, config_scene
from oracle import gags_oracle as O

cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 150
n, H, W, D = CONFIGS[cfg]
scene = config_scene(cfg)
cam = scene.cameras[0]
K = O.intrinsics_from_fov(cam.FoVx, cam.FoVy, W, H)
scales, quats, opac = O.activate(scene.scaling, scene.rotation, scene.opacity)
radii, m2d, dep, con = O.project(scene.xyz, quats, scales, cam.world_view_transform.T, K, W, H)
tw, th = (W + 15) // 16, (H + 15) // 16
cnt, keys, ids = O.isect_tiles(m2d, radii, dep, tw, th)
offs = O.isect_offsets(keys, tw * th)
o2 = torch.cat([offs, torch.tensor([keys.numel()], dtype=torch.int32)]).tolist()
op = opac.squeeze(-1)
print(f"N={n} visible={(radii>0).sum().item()} n_isects={keys.numel()} per-tile mean={keys.numel()/(tw*th):.0f} max={max(o2[i+1]-o2[i] for i in range(tw*th))}")
g = torch.Generator().manual_seed(0)
sample = torch.randperm(tw * th, generator=g)[:nt].tolist()
acc = dict(L=[], S=[], stop_scan=[], stop_surv=[], keff=[], blk=[], S_proc_nz=[])
for t in sample:
    s, e = o2[t], o2[t + 1]
    ty, tx = divmod(t, tw)
    sel = ids[s:e].long()
    mm, cc, oo = m2d[sel], con[sel], op[sel]
    for half in range(2):
        y0 = ty * 16 + half * 8
        if y0 >= H: continue
        ys = torch.arange(y0, min(y0 + 8, H)); xs = torch.arange(tx * 16, min(tx * 16 + 16, W))
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        px = torch.stack([gx.reshape(-1).float() + 0.5, gy.reshape(-1).float() + 0.5], -1)
        if e <= s:
            acc["L"].append(0); acc["S"].append(0); acc["stop_scan"].append(0); acc["stop_surv"].append(0); continue
        # bbox cull as in tc_alpha_extent
        Lg = torch.log(255.0 * oo)
        det = cc[:, 0] * cc[:, 2] - cc[:, 1] ** 2
        Lm = Lg.clamp_min(0) + 2e-3
        hx = torch.sqrt(2 * Lm / det * cc[:, 2]) * 1.0005 + 0.02
        hy = torch.sqrt(2 * Lm / det * cc[:, 0]) * 1.0005 + 0.02
        x_lo, x_hi = tx * 16 + 0.5, tx * 16 + 15.5
        y_lo, y_hi = y0 + 0.5, y0 + 7.5
        keep = (Lg > -1e-3) & (mm[:, 0] + hx >= x_lo) & (mm[:, 0] - hx <= x_hi) & (mm[:, 1] + hy >= y_lo) & (mm[:, 1] - hy <= y_hi)
        w, kp, tfin = O._tile_weights(px, mm, cc, oo)      # [P, G]
        contrib = w > 0
        acc["keff"].append(contrib.sum(1).float().mean().item())
        anyc = contrib.any(0)
        last = int(anyc.nonzero().max()) + 1 if anyc.any() else 0
        acc["L"].append(e - s); acc["S"].append(int(keep.sum()))
        acc["stop_scan"].append(last); acc["stop_surv"].append(int(keep[:last].sum()))
        acc["S_proc_nz"].append(int((anyc & keep)[:last].sum()))
        # 8x4 block occupancy among processed survivors
        if last > 0:
            cb = contrib[:, :last].reshape(len(ys), len(xs), last)
            nb = 0; tot = 0
            for by in range(0, len(ys), 4):
                for bx in range(0, len(xs), 8):
                    nb += cb[by:by+4, bx:bx+8].reshape(-1, last).any(0)[keep[:last]].sum().item(); tot += int(keep[:last].sum())
            acc["blk"].append(nb / max(tot, 1))
def m(x): return sum(x) / max(len(x), 1)
print(f"half tiles sampled {len(acc['L'])}: list {m(acc['L']):.0f}  bbox-survivors {m(acc['S']):.0f}  scanned-until-stop {m(acc['stop_scan']):.0f} "
      f"survivors-until-stop {m(acc['stop_surv']):.0f} (of which contributing {m(acc['S_proc_nz']):.0f})  batches/half-tile {m([math.ceil(x/32) for x in acc['stop_surv']]):.1f}  "
      f"K_eff/pixel {m(acc['keff']):.1f}  block occupancy {m(acc['blk']):.2f}")

# ---- how much would an exact ellipse-vs-rectangle test cull beyond the bounding box? -----------------
def min_sigma_rect(mm, cc, x_lo, x_hi, y_lo, y_hi):
    """min over the rectangle of sigma = .5(a dx^2 + c dy^2) + b dx dy (convex quadratic)."""
    a, b, c = cc[:, 0], cc[:, 1], cc[:, 2]
    mx, my = mm[:, 0], mm[:, 1]
    inside = (mx >= x_lo) & (mx <= x_hi) & (my >= y_lo) & (my <= y_hi)
    best = torch.full_like(a, float("inf"))
    for xe in (x_lo, x_hi):                       # vertical edges: x fixed, minimise over y
        dx = xe - mx
        dy = (-b * dx / c).clamp(y_lo - my, y_hi - my)
        best = torch.minimum(best, 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy)
    for ye in (y_lo, y_hi):
        dy = ye - my
        dx = (-b * dy / a).clamp(x_lo - mx, x_hi - mx)
        best = torch.minimum(best, 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy)
    return torch.where(inside, torch.zeros_like(best), best)

tot_bbox = tot_exact = 0
for t in sample[:60]:
    s, e = o2[t], o2[t + 1]
    if e <= s: continue
    ty, tx = divmod(t, tw)
    sel = ids[s:e].long()
    mm, cc, oo = m2d[sel], con[sel], op[sel]
    for half in range(2):
        y0 = ty * 16 + half * 8
        if y0 >= H: continue
        Lg = torch.log(255.0 * oo)
        det = cc[:, 0] * cc[:, 2] - cc[:, 1] ** 2
        Lm = Lg.clamp_min(0) + 2e-3
        hx = torch.sqrt(2 * Lm / det * cc[:, 2]) * 1.0005 + 0.02
        hy = torch.sqrt(2 * Lm / det * cc[:, 0]) * 1.0005 + 0.02
        x_lo, x_hi, y_lo, y_hi = tx * 16 + 0.5, tx * 16 + 15.5, y0 + 0.5, y0 + 7.5
        keep = (Lg > -1e-3) & (mm[:, 0] + hx >= x_lo) & (mm[:, 0] - hx <= x_hi) & (mm[:, 1] + hy >= y_lo) & (mm[:, 1] - hy <= y_hi)
        exact = keep & (min_sigma_rect(mm, cc, x_lo, x_hi, y_lo, y_hi) <= Lm)
        tot_bbox += int(keep.sum()); tot_exact += int(exact.sum())
print(f"exact ellipse-vs-rectangle cull keeps {tot_exact} of {tot_bbox} bbox survivors ({100*tot_exact/max(tot_bbox,1):.1f} %)")
