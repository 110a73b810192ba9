"""CPU restatement of the reference's distillation target and loss.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/scene/dataset_readers.py:54-121 (read_sam_clip_feature, default and
max_mode branches) and the call site /root/reference/train.py:162-163
(`Ll1_feature = l1_loss(feature_map * seg_mask, gt_feature_map * seg_mask)`, l1_loss =
/root/reference/utils/loss_utils.py:20-21), statement by statement, in plain PyTorch.  The reference
module itself cannot be imported here (it pulls in plyfile / PIL scene loaders), so parity of this
file is by inspection; what it pins is the product's fused kernels against the reference's maths.
"""
import torch
import torch.nn.functional as F


def read_sam_clip_feature(img_embed, seg_map, scale_map, max_mode=False):
    _, h, w = seg_map.shape                                        # :55
    _, h_scale, w_scale = scale_map.shape                          # :59
    x, y = torch.meshgrid(torch.arange(0, h), torch.arange(0, w), indexing="ij")   # :61-63
    x, y = x.reshape(-1, 1), y.reshape(-1, 1)
    seg_s = seg_map[1][x, y].squeeze(-1).long()                    # :64-66
    seg_m = seg_map[2][x, y].squeeze(-1).long()
    seg_l = seg_map[3][x, y].squeeze(-1).long()
    mask_s = (seg_s != -1).reshape(1, h, w)                        # :68-71
    mask_m = (seg_m != -1).reshape(1, h, w)
    mask_l = (seg_l != -1).reshape(1, h, w)
    mask = mask_s & mask_m & mask_l
    mask = F.interpolate(mask.float().unsqueeze(0), size=(h_scale, w_scale),
                         mode="nearest").squeeze(0).to(torch.bool)  # :72
    fs = img_embed[seg_s].reshape(h, w, -1).permute(2, 0, 1)       # :73-75 (id -1 -> last row)
    fm = img_embed[seg_m].reshape(h, w, -1).permute(2, 0, 1)
    fl = img_embed[seg_l].reshape(h, w, -1).permute(2, 0, 1)
    up = lambda t: F.interpolate(t.unsqueeze(0), size=(h_scale, w_scale), mode="bilinear",
                                 align_corners=True).squeeze(0)     # :77-79
    fs, fm, fl = up(fs), up(fm), up(fl)
    if max_mode:                                                    # :81-88
        near = lambda m_: F.interpolate(m_.float().unsqueeze(0), size=(h_scale, w_scale),
                                        mode="nearest").squeeze(0).to(torch.bool)
        mask_s, mask_m, mask_l = near(mask_s), near(mask_m), near(mask_l)
        max_idx = torch.argmax(scale_map, dim=0)
        one = F.one_hot(max_idx, num_classes=3).permute(2, 0, 1).to(scale_map.dtype)
        fmap = fs * one[0] * mask_s + fm * one[1] * mask_m + fl * one[2] * mask_l
        return fmap, fmap[0:1] != 0.0
    return fs * scale_map[0] + fm * scale_map[1] + fl * scale_map[2], mask   # :119-120


def distill_loss(feature_map, img_embed, seg_map, scale_map):
    """train.py:162-163 (iteration < scale_balance_iteration branch)."""
    gt, seg_mask = read_sam_clip_feature(img_embed, seg_map, scale_map)
    return torch.abs(feature_map * seg_mask - gt * seg_mask).mean()
