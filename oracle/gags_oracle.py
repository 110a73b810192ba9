"""CPU oracle for the GAGS feature-rasterisation hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain PyTorch tensor ops that run on the CPU in fp32 or fp64, the
algorithm the reference reaches through its single call to ``gsplat.rasterization``
(/root/reference/gaussian_renderer/__init__.py:56-70).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package ``gags_b200`` never does.

PARITY STATUS: **parity unpinned** for the rasteriser arithmetic.  ``gsplat`` is an
un-vendored, un-pinned pip dependency (/root/reference/environment.yml:26; API usage implies
1.0 <= version < 1.5, most likely 1.4.0) whose source is absent from /root/reference and
from this image, and the reference ships no tests or golden vectors.  The stages below
therefore follow the published gsplat v1.4 algorithm as written down in SURVEY.md
Appendix A.  What *is* pinned against code that exists in /root/reference (see
oracle/make_golden.py and tests/golden/):
  * quaternion -> rotation and Sigma = (R S)(R S)^T   utils/general_utils.py:78-110,
                                                      scene/gaussian_model.py:28-32
  * the SH basis                                      utils/sh_utils.py:57-112
  * focal length from FoV / K construction            gaussian_renderer/__init__.py:27-38,
                                                      utils/graphics_utils.py:73-74
  * world->view matrix                                utils/graphics_utils.py:38-49
  * pinhole projection of centres                     depth_SAM.py:34-77

Every function takes torch tensors of one floating dtype (float32 or float64) on the CPU
and is differentiable through ``torch.autograd`` wherever the path is differentiable, so
the same code gives the backward oracle (Appendix A.6/A.7).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch

TILE = 16
ALPHA_MIN = 1.0 / 255.0
ALPHA_MAX = 0.999
T_STOP = 1e-4

# SH constants, identical to /root/reference/utils/sh_utils.py:26-55
C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005,
      -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
      -0.4570457994644658, 1.445305721320277, -0.5900435899266435]
C4 = [2.5033429417967046, -1.7701307697799304, 0.9461746957575601, -0.6690465435572892,
      0.10578554691520431, -0.6690465435572892, 0.47308734787878004, -1.7701307697799304,
      0.6258357354491761]


# --------------------------------------------------------------------------------------
# camera helpers  (gaussian_renderer/__init__.py:27-38, utils/graphics_utils.py:38-49,73)
# --------------------------------------------------------------------------------------
def intrinsics_from_fov(fovx: float, fovy: float, width: int, height: int,
                        dtype=torch.float32) -> torch.Tensor:
    """K exactly as render() builds it: fx = W / (2 tan(FoVx/2)), cx = W/2."""
    fx = width / (2.0 * math.tan(fovx * 0.5))
    fy = height / (2.0 * math.tan(fovy * 0.5))
    return torch.tensor([[fx, 0.0, width / 2.0], [0.0, fy, height / 2.0], [0.0, 0.0, 1.0]],
                        dtype=dtype)


def world2view(R, t, translate=(0.0, 0.0, 0.0), scale=1.0):
    """getWorld2View2 (utils/graphics_utils.py:38-49) restated with torch (float64 inside)."""
    R = torch.as_tensor(R, dtype=torch.float64)
    t = torch.as_tensor(t, dtype=torch.float64)
    Rt = torch.zeros(4, 4, dtype=torch.float64)
    Rt[:3, :3] = R.T
    Rt[:3, 3] = t
    Rt[3, 3] = 1.0
    C2W = torch.linalg.inv(Rt)
    C2W[:3, 3] = (C2W[:3, 3] + torch.as_tensor(translate, dtype=torch.float64)) * scale
    return torch.linalg.inv(C2W).to(torch.float32)


# --------------------------------------------------------------------------------------
# activations  (scene/gaussian_model.py:34-42,116-139)
# --------------------------------------------------------------------------------------
def activate(scaling_raw, rotation_raw, opacity_raw, scaling_modifier: float = 1.0):
    scales = torch.exp(scaling_raw) * scaling_modifier
    quats = torch.nn.functional.normalize(rotation_raw)
    opac = torch.sigmoid(opacity_raw)
    return scales, quats, opac


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """(w,x,y,z) -> R, normalising first.  == utils/general_utils.py:78-99."""
    q = q / q.norm(dim=-1, keepdim=True)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([
        1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
        2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
        2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def covariance3d(quats: torch.Tensor, scales: torch.Tensor) -> torch.Tensor:
    """Sigma = (R S)(R S)^T.  == build_scaling_rotation, general_utils.py:101-110 and
    gaussian_model.py:28-32."""
    M = quat_to_rotmat(quats) * scales[..., None, :]
    return M @ M.transpose(-1, -2)


# --------------------------------------------------------------------------------------
# A.1 projection + cull
# --------------------------------------------------------------------------------------
def project(means, quats, scales, viewmat, K, width: int, height: int, eps2d: float = 0.3,
            near: float = 0.01, far: float = 1e10, radius_clip: float = 0.0):
    """SURVEY Appendix A.1.  Returns radii int32[N], means2d [N,2], depths [N], conics [N,3].

    Culled Gaussians get radii = 0 and zeros in the float outputs.
    """
    dt = means.dtype
    R = viewmat[:3, :3].to(dt)
    t = viewmat[:3, 3].to(dt)
    K = K.to(dt)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    mc = means @ R.T + t
    x, y, z = mc.unbind(-1)
    in_z = (z >= near) & (z <= far)
    zs = torch.where(in_z, z, torch.ones_like(z))          # keep the maths finite when culled
    cov = covariance3d(quats, scales)
    covc = R @ cov @ R.T
    tanx = 0.5 * width / fx
    tany = 0.5 * height / fy
    lim_xp = (width - cx) / fx + 0.3 * tanx
    lim_xn = cx / fx + 0.3 * tanx
    lim_yp = (height - cy) / fy + 0.3 * tany
    lim_yn = cy / fy + 0.3 * tany
    rz = 1.0 / zs
    tx = zs * torch.minimum(lim_xp, torch.maximum(-lim_xn, x * rz))
    ty = zs * torch.minimum(lim_yp, torch.maximum(-lim_yn, y * rz))
    zero = torch.zeros_like(rz)
    J = torch.stack([fx * rz, zero, -fx * tx * rz * rz,
                     zero, fy * rz, -fy * ty * rz * rz], dim=-1).reshape(-1, 2, 3)
    cov2d = J @ covc @ J.transpose(-1, -2)
    m2d = torch.stack([fx * x * rz + cx, fy * y * rz + cy], dim=-1)
    c00 = cov2d[:, 0, 0] + eps2d
    c01 = cov2d[:, 0, 1]
    c11 = cov2d[:, 1, 1] + eps2d
    det = c00 * c11 - c01 * c01
    ok = in_z & (det > 0)
    dets = torch.where(ok, det, torch.ones_like(det))
    conics = torch.stack([c11 / dets, -c01 / dets, c00 / dets], dim=-1)
    b = 0.5 * (c00 + c11)
    v1 = b + torch.sqrt(torch.clamp_min(b * b - det, 0.01))
    radius = torch.ceil(3.0 * torch.sqrt(torch.clamp_min(v1, 0.0)))
    ok = ok & (radius > radius_clip)
    ok = ok & ~((m2d[:, 0] + radius <= 0) | (m2d[:, 0] - radius >= width)
                | (m2d[:, 1] + radius <= 0) | (m2d[:, 1] - radius >= height))
    radii = torch.where(ok, radius, torch.zeros_like(radius)).to(torch.int32)
    okf = ok.to(dt)
    return radii, m2d * okf[:, None], z * okf, conics * okf[:, None]


# --------------------------------------------------------------------------------------
# A.2 colours from spherical harmonics  (basis == utils/sh_utils.py:57-112)
# --------------------------------------------------------------------------------------
def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """[N, (deg+1)^2] real SH basis at unit directions, sign convention of sh_utils.eval_sh."""
    x, y, z = dirs.unbind(-1)
    out = [torch.full_like(x, C0)]
    if deg > 0:
        out += [-C1 * y, C1 * z, -C1 * x]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        out += [C2[0] * xy, C2[1] * yz, C2[2] * (2.0 * zz - xx - yy), C2[3] * xz,
                C2[4] * (xx - yy)]
    if deg > 2:
        out += [C3[0] * y * (3 * xx - yy), C3[1] * xy * z, C3[2] * y * (4 * zz - xx - yy),
                C3[3] * z * (2 * zz - 3 * xx - 3 * yy), C3[4] * x * (4 * zz - xx - yy),
                C3[5] * z * (xx - yy), C3[6] * x * (xx - 3 * yy)]
    if deg > 3:
        out += [C4[0] * xy * (xx - yy), C4[1] * yz * (3 * xx - yy), C4[2] * xy * (7 * zz - 1),
                C4[3] * yz * (7 * zz - 3), C4[4] * (zz * (35 * zz - 30) + 3),
                C4[5] * xz * (7 * zz - 3), C4[6] * (xx - yy) * (7 * zz - 1),
                C4[7] * xz * (xx - 3 * yy),
                C4[8] * (xx * (xx - 3 * yy) - yy * (3 * xx - yy))]
    return torch.stack(out, dim=-1)


def sh_colors(deg: int, means, viewmat, coeffs, radii) -> torch.Tensor:
    """Appendix A.2: dir = mu - cam_pos (normalised), colours = max(SH + 0.5, 0); Gaussians
    with radii <= 0 give 0 before the +0.5 (i.e. 0.5 after), as gsplat masks SH only.
    coeffs: [N, K, 3] with K >= (deg+1)^2, layout of GaussianModel.get_features
    (scene/gaussian_model.py:127-131)."""
    dt = means.dtype
    campos = torch.linalg.inv(viewmat.to(torch.float64))[:3, 3].to(dt)
    d = means - campos
    d = d / d.norm(dim=-1, keepdim=True).clamp_min(1e-20)
    B = sh_basis(deg, d)                                        # [N, nb]
    nb = (deg + 1) ** 2
    rgb = (B[:, :, None] * coeffs[:, :nb, :]).sum(dim=1)        # [N, 3]
    rgb = rgb * (radii > 0).to(dt)[:, None]
    return torch.clamp_min(rgb + 0.5, 0.0)


# --------------------------------------------------------------------------------------
# A.3 tile intersection + keys, A.4 offsets   (integer stages: exact)
# --------------------------------------------------------------------------------------
def tile_bounds(means2d, radii, tile_w: int, tile_h: int, tile: int = TILE):
    """fp32 arithmetic exactly as A.3 (the division by 16 is exact; only the +/- rounds)."""
    m = means2d.to(torch.float32)
    r = radii.to(torch.float32) / tile
    tx = m[:, 0] / tile
    ty = m[:, 1] / tile
    x0 = torch.floor(tx - r).clamp(0, tile_w).to(torch.int64)
    x1 = torch.ceil(tx + r).clamp(0, tile_w).to(torch.int64)
    y0 = torch.floor(ty - r).clamp(0, tile_h).to(torch.int64)
    y1 = torch.ceil(ty + r).clamp(0, tile_h).to(torch.int64)
    vis = radii > 0
    z = torch.zeros_like(x0)
    return (torch.where(vis, x0, z), torch.where(vis, x1, z),
            torch.where(vis, y0, z), torch.where(vis, y1, z))


def tile_bits(n_tiles: int) -> int:
    return int(math.floor(math.log2(n_tiles))) + 1 if n_tiles > 0 else 1


def isect_tiles(means2d, radii, depths, tile_w: int, tile_h: int, tile: int = TILE,
                sort: bool = True):
    """Appendix A.3 with C = 1 (cam id 0).  Returns tiles_per_gauss int32[N],
    isect_ids int64[n_isects] (tile << 32 | depth bits), flatten_ids int32[n_isects]."""
    x0, x1, y0, y1 = tile_bounds(means2d, radii, tile_w, tile_h, tile)
    nx = x1 - x0
    cnt = nx * (y1 - y0)
    n = int(cnt.sum())
    gid = torch.repeat_interleave(torch.arange(cnt.numel()), cnt)
    start = torch.cumsum(cnt, 0) - cnt
    local = torch.arange(n) - start[gid]
    ty = y0[gid] + local // nx[gid].clamp_min(1)
    tx = x0[gid] + local % nx[gid].clamp_min(1)
    dbits = depths.to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    keys = ((ty * tile_w + tx) << 32) | dbits[gid]
    vals = gid.to(torch.int32)
    if sort:
        # emission order is ascending Gaussian index then row-major tile, and the radix
        # sort is stable -> ties on (tile, depth) resolve by Gaussian index.
        order = torch.sort(keys, stable=True).indices
        keys, vals = keys[order], vals[order]
    return cnt.to(torch.int32), keys, vals


def isect_offsets(isect_ids, n_tiles: int):
    """Appendix A.4: first sorted index whose tile >= t; trailing empties = n_isects."""
    tiles = (isect_ids >> 32).contiguous()
    return torch.searchsorted(tiles, torch.arange(n_tiles, dtype=torch.int64)).to(torch.int32)


# --------------------------------------------------------------------------------------
# A.5 blend forward (differentiable -> A.6 through autograd)
# --------------------------------------------------------------------------------------
def _tile_weights(px, m, con, op):
    """px [P,2] pixel centres; m [G,2]; con [G,3]; op [G] -> (w [P,G], keep [P,G], Tfinal [P])."""
    dx = m[None, :, 0] - px[:, None, 0]
    dy = m[None, :, 1] - px[:, None, 1]
    sigma = 0.5 * (con[None, :, 0] * dx * dx + con[None, :, 2] * dy * dy) + con[None, :, 1] * dx * dy
    vis = torch.exp(-sigma)
    alpha = torch.clamp_max(op[None, :] * vis, ALPHA_MAX)
    valid = (sigma >= 0) & (alpha >= ALPHA_MIN)
    a_eff = torch.where(valid, alpha, torch.zeros_like(alpha)).detach()
    t_incl = torch.cumprod(1.0 - a_eff, dim=1)
    stopped = valid & (t_incl <= T_STOP)
    keep = valid & ~(torch.cumsum(stopped.to(torch.int32), dim=1) > 0)
    a_k = torch.where(keep, alpha, torch.zeros_like(alpha))
    t_all = torch.cumprod(1.0 - a_k, dim=1)
    t_excl = torch.cat([torch.ones_like(t_all[:, :1]), t_all[:, :-1]], dim=1)
    return a_k * t_excl, keep, t_all[:, -1] if t_all.shape[1] else torch.ones_like(px[:, 0])


def blend_fwd(means2d, conics, colors, opacities, background, width: int, height: int,
              offsets, flatten_ids, tile: int = TILE, stats: Optional[dict] = None):
    """Appendix A.5.  Returns render [H,W,D], alpha [H,W], last_ids int32 [H,W]."""
    dt = means2d.dtype
    D = colors.shape[1]
    tile_w = (width + tile - 1) // tile
    tile_h = (height + tile - 1) // tile
    n_isects = flatten_ids.numel()
    out_rows, alpha_rows, last_rows = [], [], []
    offs = offsets.reshape(-1).tolist() + [n_isects]
    if background is None:
        background = torch.zeros(D, dtype=dt)
    n_proc = 0
    n_contrib = 0
    for ty in range(tile_h):
        row_o, row_a, row_l = [], [], []
        ys = torch.arange(ty * tile, min((ty + 1) * tile, height))
        for tx in range(tile_w):
            xs = torch.arange(tx * tile, min((tx + 1) * tile, width))
            t = ty * tile_w + tx
            s, e = offs[t], offs[t + 1]
            P = ys.numel() * xs.numel()
            if e <= s:
                row_o.append(background.expand(ys.numel(), xs.numel(), D))
                row_a.append(torch.zeros(ys.numel(), xs.numel(), dtype=dt))
                row_l.append(torch.zeros(ys.numel(), xs.numel(), dtype=torch.int32))
                continue
            ids = flatten_ids[s:e].long()
            gy, gx = torch.meshgrid(ys, xs, indexing="ij")
            px = torch.stack([gx.reshape(-1).to(dt) + 0.5, gy.reshape(-1).to(dt) + 0.5], dim=-1)
            w, keep, t_fin = _tile_weights(px, means2d[ids], conics[ids], opacities[ids])
            o = w @ colors[ids] + t_fin[:, None] * background[None, :]
            idx = torch.arange(s, e, dtype=torch.int32)[None, :].expand(P, -1)
            last = torch.where(keep, idx, torch.zeros_like(idx)).max(dim=1).values
            row_o.append(o.reshape(ys.numel(), xs.numel(), D))
            row_a.append((1.0 - t_fin).reshape(ys.numel(), xs.numel()))
            row_l.append(last.reshape(ys.numel(), xs.numel()))
            if stats is not None:
                n_contrib += int(keep.sum())
                anyk = keep.any(dim=0)
                n_proc += int(anyk.nonzero().max() + 1) if bool(anyk.any()) else 0
        out_rows.append(torch.cat(row_o, dim=1))
        alpha_rows.append(torch.cat(row_a, dim=1))
        last_rows.append(torch.cat(row_l, dim=1))
    if stats is not None:
        stats["k_eff"] = n_contrib / float(width * height)
        stats["processed_per_tile"] = n_proc / float(tile_w * tile_h)
    return torch.cat(out_rows, 0), torch.cat(alpha_rows, 0), torch.cat(last_rows, 0)


# --------------------------------------------------------------------------------------
# A.6 blend backward, written out sequentially (small cases; cross-checks autograd)
# --------------------------------------------------------------------------------------
def blend_bwd_sequential(means2d, conics, colors, opacities, background, width, height,
                         offsets, flatten_ids, render_alpha, last_ids, v_out, v_alpha,
                         tile: int = TILE):
    """Per-pixel back-to-front loop exactly as Appendix A.6 (pure Python: tiny inputs only).
    Returns v_means2d [N,2], v_conics [N,3], v_colors [N,D], v_opacities [N]."""
    dt = means2d.dtype
    N, D = colors.shape
    tile_w = (width + tile - 1) // tile
    n_isects = flatten_ids.numel()
    offs = offsets.reshape(-1).tolist() + [n_isects]
    v_m = torch.zeros(N, 2, dtype=dt)
    v_c = torch.zeros(N, 3, dtype=dt)
    v_col = torch.zeros(N, D, dtype=dt)
    v_o = torch.zeros(N, dtype=dt)
    bg = background if background is not None else torch.zeros(D, dtype=dt)
    for i in range(height):
        for j in range(width):
            t = (i // tile) * tile_w + (j // tile)
            s, e = offs[t], offs[t + 1]
            if e <= s:
                continue
            T_final = 1.0 - float(render_alpha[i, j])
            T = T_final
            S = torch.zeros(D, dtype=dt)
            vo = v_out[i, j]
            va = float(v_alpha[i, j])
            pxx, pyy = j + 0.5, i + 0.5
            for k in range(int(last_ids[i, j]), s - 1, -1):
                g = int(flatten_ids[k])
                dx = float(means2d[g, 0]) - pxx
                dy = float(means2d[g, 1]) - pyy
                a, b, c = (float(v) for v in conics[g])
                sigma = 0.5 * (a * dx * dx + c * dy * dy) + b * dx * dy
                vis = math.exp(-sigma)
                o = float(opacities[g])
                alpha = min(ALPHA_MAX, o * vis)
                if sigma < 0 or alpha < ALPHA_MIN:
                    continue
                ra = 1.0 / (1.0 - alpha)
                T = T * ra
                w = alpha * T
                v_col[g] += w * vo
                v_al = float(((colors[g] * T - S * ra) * vo).sum()) + T_final * ra * va \
                    - T_final * ra * float((bg * vo).sum())
                if o * vis <= ALPHA_MAX:
                    v_sig = -o * vis * v_al
                    v_c[g] += torch.tensor([0.5 * v_sig * dx * dx, v_sig * dx * dy,
                                            0.5 * v_sig * dy * dy], dtype=dt)
                    v_m[g] += torch.tensor([v_sig * (a * dx + b * dy), v_sig * (b * dx + c * dy)],
                                           dtype=dt)
                    v_o[g] += vis * v_al
                S = S + colors[g] * w
    return v_m, v_c, v_col, v_o


# --------------------------------------------------------------------------------------
# 3.3 the rasterization() orchestration   (call site gaussian_renderer/__init__.py:56-70)
# --------------------------------------------------------------------------------------
def rasterization(means, quats, scales, opacities, colors, viewmat, K, width: int,
                  height: int, background=None, sh_degree: Optional[int] = None,
                  render_mode: str = "RGB", eps2d: float = 0.3, near: float = 0.01,
                  far: float = 1e10, stats: Optional[dict] = None):
    """C = 1, packed=False restatement of gsplat.rasterization.
    Returns render [H,W,D'], alpha [H,W], info dict."""
    dt = means.dtype
    radii, means2d, depths, conics = project(means, quats, scales, viewmat.to(dt), K.to(dt),
                                             width, height, eps2d, near, far)
    if sh_degree is None:
        cols = colors
    else:
        cols = sh_colors(sh_degree, means, viewmat, colors, radii)
    bg = background
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[:, None]], dim=-1)
        if bg is not None:
            bg = torch.cat([bg, torch.zeros(1, dtype=dt)])
    elif render_mode in ("D", "ED"):
        cols = depths[:, None]
        bg = None
    tile_w = (width + TILE - 1) // TILE
    tile_h = (height + TILE - 1) // TILE
    tiles_per_gauss, isect_ids, flatten_ids = isect_tiles(means2d.detach(), radii,
                                                          depths.detach(), tile_w, tile_h)
    offsets = isect_offsets(isect_ids, tile_w * tile_h)
    render, alpha, last_ids = blend_fwd(means2d, conics, cols, opacities, bg, width, height,
                                        offsets, flatten_ids, stats=stats)
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1],
                            render[..., -1:] / alpha.clamp_min(1e-10)[..., None]], dim=-1)
    info = dict(radii=radii, means2d=means2d, depths=depths, conics=conics,
                opacities=opacities, tiles_per_gauss=tiles_per_gauss, isect_ids=isect_ids,
                flatten_ids=flatten_ids, isect_offsets=offsets.reshape(tile_h, tile_w),
                last_ids=last_ids, tile_width=tile_w, tile_height=tile_h, width=width,
                height=height, tile_size=TILE, n_cameras=1, colors=cols)
    return render, alpha, info


def render(camera, pc, bg_color, feature_mode=True, scaling_modifier=1.0, override_color=None,
           render_mode="RGB", dtype=torch.float32, stats=None):
    """gaussian_renderer/__init__.py:19-85 on the oracle.  `camera` needs FoVx, FoVy,
    image_width, image_height, world_view_transform (W2C transposed, scene/cameras.py:58);
    `pc` needs the raw GaussianModel tensors."""
    W, H = int(camera.image_width), int(camera.image_height)
    K = intrinsics_from_fov(camera.FoVx, camera.FoVy, W, H, dtype)
    cpu = lambda t: t.detach().to("cpu", dtype)
    scales, quats, opac = activate(cpu(pc._scaling), cpu(pc._rotation), cpu(pc._opacity),
                                   scaling_modifier)
    bg = cpu(bg_color)
    if feature_mode:
        colors = cpu(pc._semantic_feature)
        sh_degree = None
        bg = bg[0].repeat(colors.shape[-1])
    elif override_color is not None:
        colors, sh_degree = cpu(override_color), None
    else:
        colors = torch.cat([cpu(pc._features_dc), cpu(pc._features_rest)], dim=1)
        sh_degree = pc.active_sh_degree
    viewmat = cpu(camera.world_view_transform).T
    img, alpha, info = rasterization(cpu(pc._xyz), quats, scales, opac.squeeze(-1), colors,
                                     viewmat, K, W, H, bg, sh_degree, render_mode, stats=stats)
    return {"render": img.permute(2, 0, 1), "viewspace_points": info["means2d"][None],
            "visibility_filter": info["radii"] > 0, "radii": info["radii"], "alpha": alpha,
            "info": info}
