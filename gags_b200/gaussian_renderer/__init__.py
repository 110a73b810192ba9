"""render(): the reference-facing entry point, same signature, same four dict keys as
/root/reference/gaussian_renderer/__init__.py:19-85 — backed by the sm_100a kernels.

Differences that are invisible to callers (train.py:142, render.py:118, evaluate_iou_loc.py:258,
compute_relvancy.py:243,246,434):
  * K is never materialised on the device: fx, fy, cx, cy go to the kernel as scalars
    (the reference does torch.tensor(K, device="cuda") per call, :31-38);
  * exp / sigmoid / normalize (scene/gaussian_model.py:116-136) and `* scaling_modifier` (:42) are
    fused into the projection kernel when `pc` exposes the raw parameters;
  * any D renders in one pass (no 32-channel chunks, no torch.cat).
"""
from __future__ import annotations

import math

import torch

from .. import _C
from ..rasterization import rasterize_view


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, feature_mode=True,
           scaling_modifier=1.0, override_color=None, render_mode="RGB"):
    """Render the scene.  Background tensor (bg_color) must be on the GPU.

    `pipe` is accepted and ignored, as in the reference.  Positional order matches the reference
    (train.py:117 passes scaling_modifer in the feature_mode slot)."""
    width = int(viewpoint_camera.image_width)
    height = int(viewpoint_camera.image_height)
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    fx = viewpoint_camera.image_width / (2 * tanfovx)
    fy = viewpoint_camera.image_height / (2 * tanfovy)
    cx = viewpoint_camera.image_width / 2.0
    cy = viewpoint_camera.image_height / 2.0

    means3D = pc.get_xyz
    raw = all(hasattr(pc, a) for a in ("_scaling", "_rotation", "_opacity"))
    if raw:
        # fused activations: exp(), sigmoid(), normalize() happen inside the projection kernel
        scales, rotations, opacity = pc._scaling, pc._rotation, pc._opacity
        flags = _C.GAGS_F_LOG_SCALES | _C.GAGS_F_LOGIT_OPACITY
        smod = float(scaling_modifier)
    else:
        scales = pc.get_scaling * scaling_modifier
        rotations, opacity = pc.get_rotation, pc.get_opacity
        flags, smod = 0, 1.0

    if feature_mode:
        raw = getattr(pc, "_semantic_feature_for_render", None)
        colors = raw() if raw is not None else pc.get_semantic_feature   # [N, D]
        sh_degree = None
        bg_color = bg_color[0].repeat(colors.shape[-1])
    elif override_color is not None:
        colors = override_color                                # [N, 3]
        sh_degree = None
    else:
        colors = pc.get_features                               # [N, K, 3]
        sh_degree = pc.active_sh_degree

    viewmat = viewpoint_camera.world_view_transform.transpose(0, 1)
    render_colors, render_alphas, info = rasterize_view(
        means3D, rotations, scales, opacity.reshape(-1), colors, viewmat, fx, fy, cx, cy, width,
        height, background=bg_color, sh_degree=sh_degree, render_mode=render_mode,
        scaling_modifier=smod, flags=flags)

    rendered_image = render_colors.permute(2, 0, 1)            # [H,W,D] -> [D,H,W] (a view)
    fused = getattr(render_colors, "_gags_fused", None)        # see rasterization.fused_l1_backward
    if fused is not None:
        rendered_image._gags_fused = fused
    radii = info["radii"]
    try:
        info["means2d"].retain_grad()
    except Exception:
        pass
    return {"render": rendered_image,
            "viewspace_points": info["means2d"],
            "visibility_filter": radii > 0,
            "radii": radii}
