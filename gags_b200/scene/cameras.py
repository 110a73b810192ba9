"""Camera / MiniCam with the fields render() reads (/root/reference/scene/cameras.py:17-74).
Image loading and resizing stay outside the hot path; `image` may be None."""
from __future__ import annotations

import numpy as np
import torch
from torch import nn

from ..utils.graphics_utils import getProjectionMatrix, getWorld2View2


class Camera(nn.Module):
    def __init__(self, colmap_id, R, T, FoVx, FoVy, image, gt_alpha_mask, image_name, uid,
                 semantic_feature_height=None, semantic_feature_width=None, img_embed=None,
                 seg_map=None, trans=np.array([0.0, 0.0, 0.0]), scale=1.0, data_device="cuda",
                 image_width=None, image_height=None):
        super().__init__()
        self.uid, self.colmap_id = uid, colmap_id
        self.R, self.T = R, T
        self.FoVx, self.FoVy = FoVx, FoVy
        self.image_name = image_name
        self.semantic_feature_height = semantic_feature_height
        self.semantic_feature_width = semantic_feature_width
        self.img_embed, self.seg_map = img_embed, seg_map
        self.data_device = torch.device(data_device)
        if image is not None:
            img = image.clamp(0.0, 1.0).to(self.data_device)
            if gt_alpha_mask is not None:
                img = img * gt_alpha_mask.to(self.data_device)
            self.original_image = img
            self.image_width, self.image_height = img.shape[2], img.shape[1]
        else:
            self.original_image = None
            self.image_width, self.image_height = image_width, image_height
        self.zfar, self.znear = 100.0, 0.01
        self.trans, self.scale = trans, scale
        # W2C stored TRANSPOSED (scene/cameras.py:58); render() transposes it back (:55)
        w2c = torch.from_numpy(getWorld2View2(R, T, trans, scale))
        self.world_view_transform = w2c.transpose(0, 1).contiguous().to(self.data_device)
        self.projection_matrix = getProjectionMatrix(self.znear, self.zfar, FoVx, FoVy) \
            .transpose(0, 1).to(self.data_device)
        self.full_proj_transform = self.world_view_transform @ self.projection_matrix
        self.camera_center = torch.linalg.inv(self.world_view_transform.cpu())[3, :3] \
            .to(self.data_device)


class MiniCam:
    def __init__(self, width, height, fovy, fovx, znear, zfar, world_view_transform,
                 full_proj_transform):
        self.image_width, self.image_height = width, height
        self.FoVy, self.FoVx = fovy, fovx
        self.znear, self.zfar = znear, zfar
        self.world_view_transform = world_view_transform
        self.full_proj_transform = full_proj_transform
        self.camera_center = torch.linalg.inv(world_view_transform.cpu())[3, :3] \
            .to(world_view_transform.device)
