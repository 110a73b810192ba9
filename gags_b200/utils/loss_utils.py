"""Losses that consume the render and seed the backward, named as in
/root/reference/utils/loss_utils.py:20-30, plus the fused single-pass L1."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _C


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def _rows_ok(network_output, gt) -> bool:
    """CUDA [D,H,W] float32 pair with D % 4 == 0 whose first argument needs no gradient to `gt`."""
    return (network_output.is_cuda and gt.is_cuda and network_output.dim() == 3
            and network_output.shape == gt.shape and network_output.dtype == torch.float32
            and gt.dtype == torch.float32 and network_output.shape[0] % 4 == 0
            and not gt.requires_grad)


class _PixelLoss(torch.autograd.Function):
    """mode 0: l1_loss_map, mode 1: per-pixel cosine similarity (csrc/pixel_losses.cu)."""

    @staticmethod
    def forward(ctx, mode, a_hwd, b_hwd):
        a, b = a_hwd.contiguous(), b_hwd.contiguous()
        H, W, D = a.shape
        out = torch.empty(H, W, dtype=torch.float32, device=a.device)
        stats = torch.empty(H, W, 2, dtype=torch.float32, device=a.device) if mode == 1 else None
        _C.check(_C.lib.gags_pixel_loss_fwd(mode, _C.ptr(a), _C.ptr(b), H * W, D, _C.ptr(out),
                                            _C.ptr(stats), _C.stream_ptr()), "gags_pixel_loss_fwd")
        _C.count_launch()
        ctx.mode = mode
        ctx.save_for_backward(a, b, out, stats)
        return out

    @staticmethod
    def backward(ctx, g):
        a, b, out, stats = ctx.saved_tensors
        H, W, D = a.shape
        va = torch.empty_like(a)
        _C.check(_C.lib.gags_pixel_loss_bwd(ctx.mode, _C.ptr(a), _C.ptr(b),
                                            _C.ptr(g.contiguous().to(torch.float32)), _C.ptr(out),
                                            _C.ptr(stats), H * W, D, _C.ptr(va), _C.stream_ptr()),
                 "gags_pixel_loss_bwd")
        _C.count_launch()
        return None, va, None


def l1_loss_map(network_output, gt):
    """mean(|a - b|, dim=0) -> [H,W] (utils/loss_utils.py:23-24).  On CUDA [D,H,W] maps this is one
    row-reduction kernel over the channel-last rows (a permuted view of the raster is free; a
    [D,H,W]-contiguous target is transposed once); anything else takes the eager form."""
    if _rows_ok(network_output, gt):
        return _PixelLoss.apply(0, network_output.permute(1, 2, 0), gt.permute(1, 2, 0))
    return torch.abs(network_output - gt).mean(dim=0)


def l2_loss(network_output, gt):
    return ((network_output - gt) ** 2).mean()


def cos_loss(network_output, gt):
    """1 - mean over pixels of the channel cosine similarity (utils/loss_utils.py:29-30)."""
    if _rows_ok(network_output, gt):
        return 1 - _PixelLoss.apply(1, network_output.permute(1, 2, 0), gt.permute(1, 2, 0)).mean()
    return 1 - F.cosine_similarity(network_output, gt, dim=0).mean()


def _mask_f32(mask_hw, H: int, W: int):
    """[H,W] / [1,H,W] mask of any dtype (bool included) -> contiguous float32 [H*W], or None."""
    if mask_hw is None:
        return None
    if mask_hw.numel() != H * W:
        raise ValueError(f"mask must have H*W = {H * W} elements, got shape {tuple(mask_hw.shape)}")
    return mask_hw.to(torch.float32).contiguous().reshape(-1)


class _L1Fused(torch.autograd.Function):
    @staticmethod
    def forward(ctx, render_hwd, target_hwd, mask_hw, seg_hw=None, emb=None):
        """Dense target `target_hwd` [H,W,D], or (target_hwd=None) the compact pair seg_hw [H,W]
        int32 + emb [n_seg, D]."""
        _C.require_cuda(render_hwd, target_hwd, emb, mask_hw, seg_hw)
        if render_hwd.dim() != 3:
            raise ValueError("render must be [H,W,D]")
        r = render_hwd.contiguous()
        H, W, D = r.shape
        # the kernel reads the mask as float32 [H*W]; the reference's seg_mask is a bool [1,H,W]
        # tensor (scene/dataset_readers.py:118-121), so convert and check the size here
        m = _mask_f32(mask_hw, H, W)
        loss = torch.zeros(1, dtype=torch.float32, device=r.device)
        v = torch.empty_like(r)
        numel = float(H * W * D)
        if target_hwd is not None:
            if render_hwd.shape != target_hwd.shape:
                raise ValueError("render and target must both be [H,W,D]")
            t = target_hwd.contiguous()
            _C.check(_C.lib.gags_l1_loss_fused(_C.ptr(r), _C.ptr(t), _C.ptr(m), H * W, D,
                                               1.0 / numel, _C.ptr(loss), _C.ptr(v),
                                               _C.stream_ptr()), "gags_l1_loss_fused")
        else:
            if seg_hw.shape != (H, W) or seg_hw.dtype != torch.int32 or emb.dim() != 2 \
                    or emb.shape[1] != D or emb.dtype != torch.float32:
                raise ValueError("seg must be int32 [H,W] and emb float32 [n_seg, D]")
            sg, em = seg_hw.contiguous(), emb.contiguous()
            _C.check(_C.lib.gags_l1_loss_segmap(_C.ptr(r), _C.ptr(sg), _C.ptr(em), _C.ptr(m), H * W,
                                                D, em.shape[0], 1.0 / numel, _C.ptr(loss),
                                                _C.ptr(v), _C.stream_ptr()), "gags_l1_loss_segmap")
        _C.count_launch()
        ctx.save_for_backward(v)
        return loss[0] / numel

    @staticmethod
    def backward(ctx, g):
        # chain with the incoming grad_output on the device; when it is the usual 1.0 the stored
        # gradient is returned untouched (no 2 GB multiply pass).  The buffer is scaled in place,
        # so this node supports a single backward (no retain_graph re-entry).
        (v,) = ctx.saved_tensors
        gs = g.detach().reshape(1).to(torch.float32).contiguous()
        _C.check(_C.lib.gags_scale_inplace(_C.ptr(v), _C.ptr(gs), v.numel(), _C.stream_ptr()),
                 "gags_scale_inplace")
        _C.count_launch()
        return v, None, None, None, None


def l1_loss_fused(render_dhw, gt_hwd, mask_hw=None):
    """mean(|render*mask - gt*mask|) (train.py:162-163) in ONE pass that also produces the gradient.
    `render_dhw` is render()["render"] ([D,H,W] view of the channel-last raster); `gt_hwd` is the
    target in channel-last layout [H,W,D]; `mask_hw` an optional non-negative [H,W] mask."""
    return _L1Fused.apply(render_dhw.permute(1, 2, 0), gt_hwd, mask_hw)


def l1_loss_segmap_fused(render_dhw, seg_hw, emb, mask_hw=None):
    """l1_loss(render, emb[seg]) without materialising the dense target: `seg_hw` int32 [H,W]
    segment ids (< 0 = pixel without target, excluded), `emb` [n_seg, D] per-segment embeddings — the
    compact per-view inputs of read_sam_clip_feature (scene/dataset_readers.py:54-121), which the
    reference gathers into a [D,H,W] map every iteration before the loss (train.py:162-163).  The
    mean is over all H*W*D elements, as l1_loss does."""
    return _L1Fused.apply(render_dhw.permute(1, 2, 0), None, mask_hw, seg_hw, emb)


class _L1Sam(torch.autograd.Function):
    """loss of train.py:162-163 against read_sam_clip_feature's default-mode target, one pass."""

    @staticmethod
    def forward(ctx, render_hwd, seg3, emb, scale_map):
        _C.require_cuda(render_hwd, seg3, emb, scale_map)
        r = render_hwd.contiguous()
        H, W, D = r.shape
        if seg3.dtype != torch.int32 or tuple(seg3.shape) != (3, H, W):
            raise ValueError("seg3 must be int32 [3,H,W] (use sam_levels() on the reference's seg_map)")
        if emb.dim() != 2 or emb.shape[1] != D or emb.dtype != torch.float32:
            raise ValueError("img_embed must be float32 [n_seg, D]")
        if tuple(scale_map.shape) != (3, H, W) or scale_map.dtype != torch.float32:
            raise ValueError("scale_map must be float32 [3,H,W] at the render's size")
        want_vs = ctx.needs_input_grad[3]
        if want_vs and D % 128 != 0:
            raise ValueError("the gradient w.r.t. scale_map needs D % 128 == 0 on the fused path")
        sg, em, sc = seg3.contiguous(), emb.contiguous(), scale_map.contiguous()
        loss = torch.zeros(1, dtype=torch.float32, device=r.device)
        v = torch.empty_like(r)
        vs = torch.zeros_like(sc) if want_vs else None
        numel = float(H * W * D)
        _C.check(_C.lib.gags_l1_loss_sam(_C.ptr(r), _C.ptr(sg), _C.ptr(em), _C.ptr(sc), H * W, D,
                                         em.shape[0], 1.0 / numel, _C.ptr(loss), _C.ptr(v),
                                         _C.ptr(vs), _C.stream_ptr()), "gags_l1_loss_sam")
        _C.count_launch()
        ctx.save_for_backward(v, vs)
        return loss[0] / numel

    @staticmethod
    def backward(ctx, g):
        v, vs = ctx.saved_tensors
        gs = g.detach().reshape(1).to(torch.float32).contiguous()
        _C.check(_C.lib.gags_scale_inplace(_C.ptr(v), _C.ptr(gs), v.numel(), _C.stream_ptr()),
                 "gags_scale_inplace")
        _C.count_launch()
        return v, None, None, (vs * gs if vs is not None else None)


def sam_levels(seg_map: torch.Tensor) -> torch.Tensor:
    """The reference's per-view `seg_map` ([4,h,w], level 0 unused by the target: levels 1..3 = s, m, l,
    scene/dataset_readers.py:64-66) or an already-sliced [3,h,w] map -> contiguous int32 [3,h,w]."""
    if seg_map.dim() != 3 or seg_map.shape[0] not in (3, 4):
        raise ValueError("seg_map must be [4,h,w] or [3,h,w]")
    s = seg_map[1:4] if seg_map.shape[0] == 4 else seg_map
    return s.to(torch.int32).contiguous()


def l1_loss_sam_fused(render_dhw, seg_map, img_embed, scale_map):
    """`gt, m = read_sam_clip_feature(img_embed, seg_map, scale_map); l1_loss(render * m, gt * m)`
    (train.py:162-163, default mode) in one pass that never materialises `gt`: the kernel gathers
    the three levels' embeddings, weights them with `scale_map` and masks pixels where any level is
    -1.  Gradients flow to the render and (D % 128 == 0) to `scale_map`.  When the maps are not at the
    render's size the reference's resize is needed: the dense route below is taken instead."""
    D, H, W = render_dhw.shape
    seg3 = seg_map if (seg_map.dtype == torch.int32 and seg_map.shape[0] == 3
                       and seg_map.is_contiguous()) else sam_levels(seg_map)
    if tuple(seg3.shape[1:]) != (H, W) or tuple(scale_map.shape[1:]) != (H, W) or \
            (scale_map.requires_grad and D % 128 != 0):
        from ..scene.dataset_readers import read_sam_clip_feature
        full = seg_map if seg_map.shape[0] == 4 else torch.cat([seg_map[:1], seg_map])
        gt, m = read_sam_clip_feature(img_embed, full, scale_map)
        if scale_map.requires_grad:
            return l1_loss(render_dhw * m, gt * m)
        return l1_loss_fused(render_dhw, gt.permute(1, 2, 0).contiguous(), m)
    return _L1Sam.apply(render_dhw.permute(1, 2, 0), seg3, img_embed, scale_map)


def l1_backward_fused_sam(render_dhw, seg_map, img_embed, scale_map, want_scale_grad=False):
    """l1_loss_sam_fused + backward as ONE kernel inside the cached feature backward
    (rasterization.fused_sam_backward); returns (detached loss, d loss / d scale_map or None)."""
    from ..rasterization import fused_sam_backward
    seg3 = seg_map if (seg_map.dtype == torch.int32 and seg_map.shape[0] == 3
                       and seg_map.is_contiguous()) else sam_levels(seg_map)
    return fused_sam_backward(render_dhw, seg3, img_embed, scale_map, want_scale_grad)


def l1_backward_fused(render_dhw, seg_hw, emb, mask_hw=None):
    """`loss = l1_loss_segmap_fused(...); loss.backward()` as ONE call that never materialises the
    [H,W,D] loss gradient (rasterization.fused_l1_backward); returns the detached loss."""
    from ..rasterization import fused_l1_backward
    return fused_l1_backward(render_dhw, seg_hw, emb, mask_hw)
