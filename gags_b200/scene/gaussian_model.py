"""GaussianModel: the per-Gaussian parameter store the renderer reads, with the reference's
operator surface (/root/reference/scene/gaussian_model.py): getters :116-139, training_setup
:183-212, capture/restore :63-113 (same 12-/13-tuple layout), save_ply/load_ply :222-319 (plyfile-free, same
file layout incl. semantic_{i}), rewrite_semantic_feature, SH degree stepping.  Densify/prune is not
exercised with frozen geometry (train.py:207) and is out of scope.

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
        self._flush_lazy()
        return (self.active_sh_degree, self._xyz, self._features_dc, self._features_rest,
                self._scaling, self._rotation, self._opacity, self.max_radii2D,
                self.xyz_gradient_accum, self.denom, self.optimizer.state_dict(),
                self.spatial_lr_scale, self._semantic_feature)

    def restore(self, model_args, training_args, resume_optimizer: bool = False):
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
        # NOTE this means a resumed run restarts with zero Adam moments in the reference (and here, by
        # default).  resume_optimizer=True additionally loads the saved state into the NEW optimiser
        # (Adam and FusedAdam share the state layout); it is an extension, off by default so that a
        # resumed run matches the reference step for step.
        if len(model_args) == 13 and self.optimizer is not None:
            self.optimizer.load_state_dict(opt_dict)
        fused = self.optimizer is not None and type(self.optimizer).__name__ == "FusedAdam"
        self.training_setup(training_args, fused_optimizer=fused)
        if resume_optimizer and len(model_args) == 13:
            self.optimizer.load_state_dict(opt_dict)
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
        # a lazily-updated table (FusedAdam(lazy_rows)) is materialised for whoever reads it through
        # the reference's getter; render() takes it through _semantic_feature_for_render() instead and
        # brings only the rows the view reads up to date
        self._flush_lazy()
        return self._semantic_feature

    def _semantic_feature_for_render(self):
        return self._semantic_feature

    def rewrite_semantic_feature(self, x):
        self._semantic_feature = x

    def get_covariance(self, scaling_modifier=1):
        return self.covariance_activation(self.get_scaling, scaling_modifier, self._rotation)

    def oneupSHdegree(self):
        if self.active_sh_degree < self.max_sh_degree:
            self.active_sh_degree += 1

    # ---- PLY I/O (scene/gaussian_model.py:222-259, :266-319), without the plyfile package ---------
    def construct_list_of_attributes(self):
        names = ["x", "y", "z", "nx", "ny", "nz"]
        names += [f"f_dc_{i}" for i in range(self._features_dc.shape[1] * self._features_dc.shape[2])]
        names += [f"f_rest_{i}" for i in
                  range(self._features_rest.shape[1] * self._features_rest.shape[2])]
        names.append("opacity")
        names += [f"scale_{i}" for i in range(self._scaling.shape[1])]
        names += [f"rot_{i}" for i in range(self._rotation.shape[1])]
        if self._semantic_feature is not None:
            names += [f"semantic_{i}" for i in range(self._semantic_feature.shape[1])]
        return names

    def _flush_lazy(self):
        """A lazily-updated feature table (FusedAdam(lazy_rows)) is materialised before it is read."""
        opt = getattr(self, "optimizer", None)
        if opt is not None and hasattr(opt, "flush"):
            opt.flush()

    def save_ply(self, path):
        self._flush_lazy()
        """Same file the reference writes: one float32 `vertex` element, SH coefficients stored
        channel-major (the [N,K,3] tensors transposed to [N,3,K] and flattened), raw (pre-activation)
        opacity / scale / rotation, then semantic_{i}."""
        import numpy as np
        from ..utils.ply_io import write_vertex_ply
        cpu = lambda t: t.detach().cpu().numpy()
        xyz = cpu(self._xyz)
        cols = [xyz, np.zeros_like(xyz),
                cpu(self._features_dc.detach().transpose(1, 2).flatten(start_dim=1).contiguous()),
                cpu(self._features_rest.detach().transpose(1, 2).flatten(start_dim=1).contiguous()),
                cpu(self._opacity), cpu(self._scaling), cpu(self._rotation)]
        if self._semantic_feature is not None:
            cols.append(cpu(self._semantic_feature))
        write_vertex_ply(path, self.construct_list_of_attributes(), np.concatenate(cols, axis=1))

    def load_ply(self, path):
        """Reads a GAGS point cloud (with semantic_{i}) or a plain 3DGS one (without; the feature
        table then stays unset until training_setup() creates it)."""
        import numpy as np
        from ..utils.ply_io import read_vertex_ply
        names, col = read_vertex_ply(path)
        dev = self.device

        def stack(prefix):
            keys = sorted((n for n in names if n.startswith(prefix)),
                          key=lambda x: int(x.split("_")[-1]))
            return np.stack([col[k] for k in keys], axis=1).astype(np.float32) if keys else None

        xyz = np.stack([col["x"], col["y"], col["z"]], axis=1).astype(np.float32)
        n = xyz.shape[0]
        f_dc = stack("f_dc_")
        f_rest = stack("f_rest_")
        k_rest = (self.max_sh_degree + 1) ** 2 - 1
        if f_rest is None or f_rest.shape[1] != 3 * k_rest:
            raise ValueError(f"{path}: expected {3 * k_rest} f_rest_* properties for SH degree "
                             f"{self.max_sh_degree}")
        par = lambda a: nn.Parameter(torch.tensor(a, dtype=torch.float32, device=dev)
                                     .contiguous().requires_grad_(True))
        self._xyz = par(xyz)
        self._features_dc = par(f_dc.reshape(n, 3, 1).transpose(0, 2, 1))
        self._features_rest = par(f_rest.reshape(n, 3, k_rest).transpose(0, 2, 1))
        self._opacity = par(col["opacity"].astype(np.float32)[:, None])
        self._scaling = par(stack("scale_"))
        self._rotation = par(stack("rot_"))
        sem = stack("semantic_")
        if sem is not None:
            self._semantic_feature = par(sem)
        self.max_radii2D = torch.zeros(n, device=dev)
        self.active_sh_degree = self.max_sh_degree

    # ---- construction from tensors (stands in for create_from_pcd) -------------------------------
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
            # row-sparse gradient handling (optim.py): on unless GAGS_B200_SPARSE_ADAM=0
            import os
            sparse = os.environ.get("GAGS_B200_SPARSE_ADAM", "1") != "0"
            # ... and lazily evaluated on top of that unless GAGS_B200_LAZY_ADAM=0: between flushes
            # `_semantic_feature` holds each row as of its last visit; render(), capture() and
            # save_ply() handle that, other readers call self.optimizer.flush() first
            lazy = sparse and os.environ.get("GAGS_B200_LAZY_ADAM", "1") != "0"
            self.optimizer = FusedAdam(groups, lr=0.0, eps=1e-15, sparse_rows=sparse,
                                       lazy_rows=lazy)
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
