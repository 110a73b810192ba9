"""CPU timing of the oracle on a bounded sample of a benchmark workload.  TEST/BENCH
INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg and `--impl reference`).

gsplat — the reference's real implementation of this path — is CUDA-only and absent from this
image, so the reference has no CPU implementation to compile into oracle/_ref; the CPU number is
the oracle port (kind = "port").  One "view" = projection + tile keys + sort + offsets on ALL N
Gaussians, plus blend forward + feature backward on `n_tiles` sampled tiles, extrapolated to the
full tile grid."""
from __future__ import annotations

import os
import time

import torch

from . import gags_oracle as O


def time_view(scene, cam, D: int, n_tiles: int = 256, seed: int = 0, threads: int | None = None,
              backward: bool = True):
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    W, H = int(cam.image_width), int(cam.image_height)
    K = O.intrinsics_from_fov(cam.FoVx, cam.FoVy, W, H)
    t0 = time.perf_counter()
    scales, quats, opac = O.activate(scene.scaling, scene.rotation, scene.opacity)
    radii, m2d, dep, con = O.project(scene.xyz, quats, scales, cam.world_view_transform.T.cpu(), K,
                                     W, H)
    tw, th = (W + 15) // 16, (H + 15) // 16
    cnt, keys, ids = O.isect_tiles(m2d, radii, dep, tw, th)
    offs = O.isect_offsets(keys, tw * th)
    t_geom = time.perf_counter() - t0
    o2 = torch.cat([offs, torch.tensor([keys.numel()], dtype=torch.int32)]).tolist()
    g = torch.Generator().manual_seed(seed)
    sample = torch.randperm(tw * th, generator=g)[:n_tiles].tolist()
    feats = scene.semantic_feature[:, :D]
    op = opac.squeeze(-1)
    per_tile, keff = [], []
    for t in sample:
        tt = time.perf_counter()
        s, e = o2[t], o2[t + 1]
        if e <= s:
            continue
        ty, tx = divmod(t, tw)
        ys = torch.arange(ty * 16, min(ty * 16 + 16, H))
        xs = torch.arange(tx * 16, min(tx * 16 + 16, W))
        gy, gx = torch.meshgrid(ys, xs, indexing="ij")
        px = torch.stack([gx.reshape(-1).float() + 0.5, gy.reshape(-1).float() + 0.5], -1)
        sel = ids[s:e].long()
        w, keep, t_fin = O._tile_weights(px, m2d[sel], con[sel], op[sel])
        f = feats[sel]
        out = w @ f                                   # forward blend
        if backward:
            v_out = torch.sign(out)                   # dL1/dout
            _ = w.T @ v_out                           # feature backward
        per_tile.append((time.perf_counter() - tt) * 1e3)
        keff.append(float((w > 0).sum(dim=1).float().mean()))      # (outside the timed part)
    t_blend = sum(per_tile) * 1e-3 * (tw * th) / max(1, len(sample))
    pt = sorted(per_tile) or [0.0]
    spread = [round(pt[int(q * (len(pt) - 1))], 3) for q in (0.1, 0.5, 0.9)]
    total = t_geom + t_blend
    return dict(seconds_per_view=total, views_per_s=1.0 / total, geom_s=t_geom,
                blend_s_extrapolated=t_blend, cores=threads, tile_ms_p10_p50_p90=spread,
                sample=f"projection+keys+sort on all N={scene.xyz.shape[0]}; blend "
                       f"{'fwd+bwd_feat' if backward else 'fwd'} on {len(sample)} of {tw * th} "
                       "tiles, extrapolated",
                n_isects=int(keys.numel()), n_visible=int((radii > 0).sum()),
                # workload descriptors every result row carries
                gauss_per_tile_mean=round(keys.numel() / float(tw * th), 1),
                gauss_per_tile_max=int(max(o2[i + 1] - o2[i] for i in range(tw * th))),
                k_eff_sampled=round(sum(keff) / max(1, len(keff)), 1))
