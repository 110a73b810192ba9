"""GaussianModel: the per-Gaussian parameter store the renderer reads, with the reference's
operator surface (/root/reference/scene/gaussian_model.py): getters :116-139, training_setup
:183-212, capture/restore :63-113 (same 12-/13-tuple layout), rewrite_semantic_feature, SH degree
stepping.  Densify/prune and PLY I/O are not exercised with frozen geometry (train.py:207) and are
out of scope.

Two additions that do not change the surface:
  * tensors live on `device` (default "cuda") instead of a hard-coded "cuda";
  * `fused_optimizer=True` in training_setup() swaps torch.optim.Adam for the single-pass
    sm_100a Adam kernel while keeping the Adam state_dict layout.
"""
from __future__ import annotations

import torch
from torch import nn

from ..utils.general_utils import (build_scaling_rotation, get_expon_lr_func, inverse_sigmoid,
                                   strip_symmetric)


class GaussianModel:
    def setup_functions(self):
        def build_covariance_from_scaling_rotation(scaling, scaling_modifier, rotation):
            L = build_scaling_rotation(scaling_modifier * scaling, rotation)
            return strip_symmetric(L @ L.transpose(1, 2))

        self.scaling_activation = torch.exp
        self.scaling_inverse_activation = torch.log
        self.covariance_activation = build_covariance_from_scaling_rotation
        self.opacity_activation = torch.sigmoid
        self.inverse_opacity_activation = inverse_sigmoid
        self.rotation_activation = torch.nn.functional.normalize

    def __init__(self, sh_degree: int, device="cuda"):
        self.device = torch.device(device)
        self.active_sh_degree = 0
        self.max_sh_degree = sh_degree
        self._xyz = torch.empty(0)
        self._features_dc = torch.empty(0)
        self._features_rest = torch.empty(0)
        self._scaling = torch.empty(0)
        self._rotation = torch.empty(0)
        self._opacity = torch.empty(0)
        self.max_radii2D = torch.empty(0)
        self.xyz_gradient_accum = torch.empty(0)
        self.denom = torch.empty(0)
        self.optimizer = None
        self.percent_dense = 0
        self.spatial_lr_scale = 0
        self.setup_functions()
        self._semantic_feature = None

    # ---- state tuple (scene/gaussian_model.py:63-113) -------------------------------------------
    def capture(self):
        return (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest,
                self._scaling, self._rotation, self._opacity, self.max_radii2D,
                self.xyz_gradient_accum, self.denom, self.optimizer.state_dict(),
                self.spatial_lr_scale, self._semantic_feature)

    def restore(self, model_args, training_args):
        if len(model_args) == 13:       # resume a feature-field run
            (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest,
             self._scaling, self._rotation, self._opacity, self.max_radii2D, xyz_gradient_accum,
             denom, opt_dict, self.spatial_lr_scale, self._semantic_feature) = model_args
        elif len(model_args) == 12:     # start feature training from an RGB 3DGS checkpoint
            (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest,
             self._scaling, self._rotation, self._opacity, self.max_radii2D, xyz_gradient_accum,
             denom, opt_dict, self.spatial_lr_scale) = model_args
        else:
            raise ValueError(f"checkpoint tuple of length {len(model_args)} (expected 12 or 13)")
        self.device = self._xyz.device
        # The reference loads the optimiser state before training_setup() rebuilds the optimiser
        # (:96 then :108), i.e. on self.optimizer from the earlier training_setup(); do the same
        # when one exists, then rebuild.
        if len(model_args) == 13 and self.optimizer is not None:
            self.optimizer.load_state_dict(opt_dict)
        self.training_setup(training_args)
        self.xyz_gradient_accum = xyz_gradient_accum
        self.denom = denom

    # ---- getters (:116-139) ---------------------------------------------------------------------
    @property
    def get_scaling(self):
        return self.scaling_activation(self._scaling)

    @property
    def get_rotation(self):
        return self.rotation_activation(self._rotation)

    @property
    def get_xyz(self):
        return self._xyz

    @property
    def get_features(self):
        return torch.cat((self._features_dc, self._features_rest), dim=1)

    @property
    def get_opacity(self):
        return self.opacity_activation(self._opacity)

    @property
    def get_semantic_feature(self):
        return self._semantic_feature

    def rewrite_semantic_feature(self, x):
        self._semantic_feature = x

    def get_covariance(self, scaling_modifier=1):
        return self.covariance_activation(self.get_scaling, scaling_modifier, self._rotation)

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    # ---- construction from tensors (stands in for create_from_pcd / load_ply) -------------------
    def create_from_tensors(self, xyz, scaling, rotation, opacity, features_dc=None,
                            features_rest=None, semantic_feature=None, spatial_lr_scale=1.0):
        dev = self.device
        n = xyz.shape[0]
        k_rest = (self.max_sh_degree + 1) ** 2 - 1
        if features_dc is None:
            features_dc = torch.zeros(n, 1, 3)
        if features_rest is None:
            features_rest = torch.zeros(n, k_rest, 3)
        mk = lambda t: nn.Parameter(t.to(dev, torch.float32).contiguous().requires_grad_(True))
        self.spatial_lr_scale = spatial_lr_scale
        self._xyz = mk(xyz)
        self._features_dc = mk(features_dc)
        self._features_rest = mk(features_rest)
        self._scaling = mk(scaling)
        self._rotation = mk(rotation)
        self._opacity = mk(opacity.reshape(n, 1))
        self.max_radii2D = torch.zeros(n, device=dev)
        if semantic_feature is not None:
            self._semantic_feature = mk(semantic_feature)
        return self

    # ---- optimiser (:183-212): only the semantic feature trains, geometry is frozen -------------
    def training_setup(self, training_args, semantic_dim=16, fused_optimizer=False):
        dev = self._xyz.device
        n = self.get_xyz.shape[0]
        self.percent_dense = training_args.percent_dense
        self.xyz_gradient_accum = torch.zeros((n, 1), device=dev)
        self.denom = torch.zeros((n, 1), device=dev)
        if self._semantic_feature is None or self._semantic_feature.shape[0] != n:
            self._semantic_feature = nn.Parameter(
                torch.zeros((n, semantic_dim), dtype=torch.float32, device=dev)
                .contiguous().requires_grad_(True))
        groups = [{"params": [self._semantic_feature], "lr": training_args.semantic_feature_lr,
                   "name": "semantic_feature"}]
        for t in (self._xyz, self._features_dc, self._features_rest, self._opacity, self._scaling,
                  self._rotation):
            t.requires_grad_(False)
        if fused_optimizer:
            from ..optim import FusedAdam
            self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15)
        else:
            self.optimizer = torch.optim.Adam(groups, lr=0.0, eps=1e-15)
        self.xyz_scheduler_args = get_expon_lr_func(
            lr_init=training_args.position_lr_init * self.spatial_lr_scale,
            lr_final=training_args.position_lr_final * self.spatial_lr_scale,
            lr_delay_mult=training_args.position_lr_delay_mult,
            max_steps=training_args.position_lr_max_steps)

    def update_learning_rate(self, iteration):
        """Per-step LR schedule; only an "xyz" group is scheduled (none exists when geometry is
        frozen, so this returns None exactly like the reference, :214-220)."""
        for group in self.optimizer.param_groups:
            if group["name"] == "xyz":
                lr = self.xyz_scheduler_args(iteration)
                group["lr"] = lr
                return lr
        return None
