"""A LABELLED restatement of the gsplat-v1.4 execution plan the reference reaches through
/root/reference/gaussian_renderer/__init__.py:56-70 — NOT gsplat itself (which is CUDA-only
third-party code, absent from /root/reference and from this image).

bench.py times this next to the product (`--impl restatement`, and `cuda_baseline` in the product
line) so that BASELINE.json's "1.5x the reference CUDA rasteriser" has a CUDA-class number beside
it.  What it reproduces is gsplat's *shape of work* on the same GPU:

  * render(): activations as separate PyTorch ops (scene/gaussian_model.py:116-136), K uploaded per
    call (gaussian_renderer/__init__.py:31-38);
  * projection -> tile count -> cumsum + blocking .item() -> emit -> global CUB radix sort of the
    (int64 key, int32 id) pairs -> offsets (gsplat isect_tiles / isect_offset_encode);
  * D > channel_chunk = 32: ceil(D/32) SEPARATE 16x16-tile SIMT launches on `.contiguous()` 32-channel
    slices (blend_fwd_narrow<32> = the rasterize_to_pixels_fwd<32> algorithm: 256-Gaussian batches
    staged in shared memory, per-pixel front-to-back loop, accumulators in registers) + torch.cat;
  * backward per chunk with ALL FOUR gradients (v_colors, v_means2d, v_conics, v_opacities) —
    gsplat's kernel always forms them — by warp reduction + atomicAdd (blend_bwd_narrow<32, FULL>),
    the slices summed by autograd;
  * loss = eager PyTorch l1_loss(render * mask, gt * mask) on a dense [D,H,W] target
    (train.py:162-163, utils/loss_utils.py:20-21); torch.optim.Adam (gaussian_model.py:199).

The kernels are this repo's own SIMT kernels (csrc/blend_fwd.cu, csrc/blend_bwd.cu), written to the
same algorithm with a different data layout; they are not upstream code and the number must be read
as "gsplat-algorithm restatement (ours)", never as a gsplat measurement.
"""
from __future__ import annotations

import math

import torch

from gags_b200 import _C
from gags_b200 import rasterization as R

CHANNEL_CHUNK = 32


class _ChunkBlend(torch.autograd.Function):
    """One <= 32-channel rasterize_to_pixels call: narrow SIMT forward, full VJP backward."""

    @staticmethod
    def forward(ctx, colors, geom, offsets, flatten_ids, bg, width, height):
        colors = colors.contiguous()
        N, D = colors.shape
        dev = colors.device
        render = torch.empty(height, width, D, dtype=torch.float32, device=dev)
        alphas = torch.empty(height, width, dtype=torch.float32, device=dev)
        last_ids = torch.empty(height, width, dtype=torch.int32, device=dev)
        # pin the SIMT kernels: the product would route a 32-channel chunk to its tensor-core kernel
        prev = _C.lib.gags_get_blend_impl()
        _C.check(_C.lib.gags_set_blend_impl(1))
        try:
            rc = _C.lib.gags_blend_fwd(_C.ptr(geom), _C.ptr(colors), D, _C.ptr(bg), width, height,
                                       _C.ptr(offsets), _C.ptr(flatten_ids), _C.ptr(render),
                                       _C.ptr(alphas), _C.ptr(last_ids), _C.stream_ptr())
        finally:
            _C.lib.gags_set_blend_impl(prev)
        _C.check(rc, "gags_blend_fwd")
        _C.count_launch()
        ctx.dims = (width, height, D, N)
        ctx.save_for_backward(colors, geom, offsets, flatten_ids, bg, alphas, last_ids)
        ctx.mark_non_differentiable(alphas, last_ids)
        return render, alphas, last_ids

    @staticmethod
    def backward(ctx, v_render, _va, _vl):
        colors, geom, offsets, flatten_ids, bg, alphas, last_ids = ctx.saved_tensors
        width, height, D, N = ctx.dims
        dev = colors.device
        v_render = v_render.contiguous()
        v_colors = torch.zeros(N, D, device=dev)
        # the geometry gradients are formed (and discarded: geometry is frozen) in every chunk
        v_m = torch.zeros(N, 2, device=dev)
        v_c = torch.zeros(N, 3, device=dev)
        v_o = torch.zeros(N, device=dev)
        _C.check(_C.lib.gags_blend_bwd_full(_C.ptr(geom), _C.ptr(colors), D, _C.ptr(bg), width,
                                            height, _C.ptr(offsets), _C.ptr(flatten_ids),
                                            _C.ptr(alphas), _C.ptr(last_ids), _C.ptr(v_render), None,
                                            _C.ptr(v_m), _C.ptr(v_c), _C.ptr(v_o), _C.ptr(v_colors),
                                            _C.stream_ptr()), "gags_blend_bwd_full")
        _C.count_launch()
        return v_colors, None, None, None, None, None, None


def rasterization(means, quats, scales, opacities, colors, viewmat, K, width, height, background):
    """gsplat.rasterization(packed=False, sh_degree=None, render_mode="RGB") for one camera:
    returns (render [H,W,D], alphas [H,W], radii [N], means2d [N,2])."""
    tile_w, tile_h = (width + 15) // 16, (height + 15) // 16
    Kh = K.detach().cpu().tolist()                       # the reference builds K on the host
    cam, keep = R.make_camera(viewmat, Kh[0][0], Kh[1][1], Kh[0][2], Kh[1][2], width, height)
    with torch.no_grad():                                # frozen geometry: no projection VJP
        radii, means2d, depths, conics, opac, tiles, geom = R._Project.apply(
            means, quats, scales, opacities, cam, keep, tile_w, tile_h)
        saved = R.bucket_sort
        R.bucket_sort = False                            # count / cumsum / .item() / emit / CUB sort
        try:
            binned = R.bin_and_sort(means2d, radii, depths, tiles, tile_w, tile_h)
        finally:
            R.bucket_sort = saved
    D = colors.shape[1]
    outs, alphas = [], None
    for c0 in range(0, D, CHANNEL_CHUNK):
        c1 = min(D, c0 + CHANNEL_CHUNK)
        chunk = colors[:, c0:c1].contiguous()
        bg = background[c0:c1].contiguous() if background is not None else None
        r, a, _ = _ChunkBlend.apply(chunk, geom, binned["offsets"], binned["flatten_ids"], bg, width,
                                    height)
        outs.append(r)
        if alphas is None:
            alphas = a
    render = torch.cat(outs, dim=-1) if len(outs) > 1 else outs[0]
    return render, alphas, radii, means2d


def render(viewpoint_camera, pc, pipe, bg_color, feature_mode=True, scaling_modifier=1.0):
    """/root/reference/gaussian_renderer/__init__.py:19-85, feature mode, on the restatement."""
    W, H = int(viewpoint_camera.image_width), int(viewpoint_camera.image_height)
    tanfovx = math.tan(viewpoint_camera.FoVx * 0.5)
    tanfovy = math.tan(viewpoint_camera.FoVy * 0.5)
    fx, fy = W / (2 * tanfovx), H / (2 * tanfovy)
    dev = pc.get_xyz.device
    K = torch.tensor([[fx, 0, W / 2.0], [0, fy, H / 2.0], [0, 0, 1.0]], device=dev)   # :31-38
    means3D = pc.get_xyz
    opacity = pc.get_opacity                             # sigmoid        (separate kernels, :40-43)
    scales = pc.get_scaling * scaling_modifier           # exp, mul
    rotations = pc.get_rotation                          # normalize
    colors = pc.get_semantic_feature
    bg = bg_color[0].repeat(colors.shape[-1])
    viewmat = viewpoint_camera.world_view_transform.transpose(0, 1)
    img, alphas, radii, means2d = rasterization(means3D, rotations, scales, opacity.reshape(-1),
                                                colors, viewmat, K, W, H, bg)
    return {"render": img.permute(2, 0, 1), "viewspace_points": means2d[None],
            "visibility_filter": radii > 0, "radii": radii}


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()          # utils/loss_utils.py:20-21
