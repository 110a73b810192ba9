"""read_sam_clip_feature: the distillation target of the training loop, restated from
/root/reference/scene/dataset_readers.py:54-121 (called at /root/reference/train.py:162-166) with
plain PyTorch ops on whatever device its inputs live on.

This is the DENSE route — it materialises the [C,h,w] target like the reference does — and serves
(i) as the fallback of utils.loss_utils.l1_loss_sam_fused when the maps have to be resized and
(ii) as the statement the fused kernels (gags_l1_loss_sam, gags_blend_bwd_features_cached_sam) are
tested against.  The COLMAP / Blender scene readers of the same reference file are data I/O and
out of scope.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _resize(x: torch.Tensor, size, mode: str) -> torch.Tensor:
    if tuple(x.shape[-2:]) == tuple(size):
        return x                      # align_corners=True bilinear / nearest at equal size = identity
    kw = {"align_corners": True} if mode == "bilinear" else {}
    return F.interpolate(x.unsqueeze(0), size=size, mode=mode, **kw).squeeze(0)


def read_sam_clip_feature(img_embed, seg_map, scale_map, max_mode: bool = False,
                          median_mode: bool = False):
    """img_embed [n_seg, C]; seg_map [4,h,w] (levels 1..3 = s, m, l; -1 = no segment); scale_map
    [3,hs,ws] level weights.  Returns (feature_map [C,hs,ws], mask bool [1,hs,ws])."""
    _, h, w = seg_map.shape
    size = tuple(scale_map.shape[-2:])
    levels = [seg_map[i].long() for i in (1, 2, 3)]
    valid = [(s != -1) for s in levels]
    mask = _resize((valid[0] & valid[1] & valid[2]).reshape(1, h, w).float(), size, "nearest").bool()
    # an id of -1 indexes the LAST embedding row, as in the reference (the mask removes it)
    feats = [_resize(img_embed[s.reshape(-1)].reshape(h, w, -1).permute(2, 0, 1), size, "bilinear")
             for s in levels]
    if max_mode:
        vm = [_resize(v.reshape(1, h, w).float(), size, "nearest").bool() for v in valid]
        one_hot = F.one_hot(torch.argmax(scale_map, dim=0), num_classes=3).permute(2, 0, 1).float()
        fmap = sum(f * one_hot[i] * vm[i] for i, f in enumerate(feats))
        return fmap, fmap[0:1] != 0.0
    if median_mode:
        seg_r = _resize(seg_map.float(), size, "nearest")
        ref_level = seg_r[1]
        weights = scale_map.clone()
        ids = ref_level[ref_level != -1]
        if ids.numel():
            for i in range(int(ids.min()), int(ref_level.max()) + 1):
                sel = ref_level == i
                if bool(sel.any()):
                    med = torch.median(scale_map[:, sel], dim=1)[0]
                    weights[:, sel] = (med / med.sum()).unsqueeze(-1)
        return sum(f * weights[i] for i, f in enumerate(feats)), mask
    return sum(f * scale_map[i] for i, f in enumerate(feats)), mask
