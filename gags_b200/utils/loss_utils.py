"""Losses that consume the render and seed the backward, named as in
/root/reference/utils/loss_utils.py:20-30, plus the fused single-pass L1."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import _C


def l1_loss(network_output, gt):
    return torch.abs(network_output - gt).mean()


def l1_loss_map(network_output, gt):
    return torch.abs(network_output - gt).mean(dim=0)


def l2_loss(network_output, gt):
    return ((network_output - gt) ** 2).mean()


def cos_loss(network_output, gt):
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


def l1_backward_fused(render_dhw, seg_hw, emb, mask_hw=None):
    """`loss = l1_loss_segmap_fused(...); loss.backward()` as ONE call that never materialises the
    [H,W,D] loss gradient (rasterization.fused_l1_backward); returns the detached loss."""
    from ..rasterization import fused_l1_backward
    return fused_l1_backward(render_dhw, seg_hw, emb, mask_hw)
