"""Seeded synthetic scenes and cameras (SURVEY.md §8d / Appendix B — normative for bench.py).

There is no dataset or checkpoint in this environment, so every benchmark and parity test
runs on Gaussians drawn here.  The generator is CPU/torch only and deterministic per seed;
raw (pre-activation) parameters are produced, as GaussianModel stores them
(/root/reference/scene/gaussian_model.py:116-139), so render() pays the activation cost.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import torch

CONFIGS = {
    # id: (N, H, W, D)   BASELINE.json "configs"
    1: (10_000, 256, 256, 3),
    2: (500_000, 720, 1280, 32),
    3: (2_000_000, 1080, 1920, 256),
    4: (2_000_000, 1080, 1920, 256),
    5: (5_000_000, 1440, 2560, 512),
}


class SynthCamera:
    """Duck-type of scene/cameras.py:Camera / MiniCam — only the fields render() reads
    (gaussian_renderer/__init__.py:27-38,55)."""

    def __init__(self, uid: int, w2c: torch.Tensor, fovx: float, fovy: float, width: int,
                 height: int):
        self.uid = uid
        self.FoVx = fovx
        self.FoVy = fovy
        self.image_width = width
        self.image_height = height
        self.znear = 0.01
        self.zfar = 100.0
        # stored TRANSPOSED exactly like scene/cameras.py:58
        self.world_view_transform = w2c.T.contiguous()
        self.camera_center = torch.linalg.inv(w2c.double())[:3, 3].float()

    def to(self, device):
        self.world_view_transform = self.world_view_transform.to(device)
        self.camera_center = self.camera_center.to(device)
        return self


def ring_cameras(n_views: int, width: int, height: int, fovx_deg: float = 60.0,
                 ring_radius: float = 11.0) -> List[SynthCamera]:
    """`n_views` pinhole cameras on a circle in the xz-plane looking at the origin.
    Camera frame: +x right, +y down, +z forward (COLMAP/3DGS convention)."""
    fovx = math.radians(fovx_deg)
    fx = width / (2.0 * math.tan(fovx / 2.0))
    fovy = 2.0 * math.atan(height / (2.0 * fx))          # square pixels
    cams = []
    for v in range(n_views):
        th = 2.0 * math.pi * v / n_views
        c = torch.tensor([ring_radius * math.sin(th), 0.0, -ring_radius * math.cos(th)],
                         dtype=torch.float64)
        fwd = -c / c.norm()
        up = torch.tensor([0.0, -1.0, 0.0], dtype=torch.float64)   # world up = -y (y is down)
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        down = torch.linalg.cross(fwd, right)
        Rwc = torch.stack([right, down, fwd], dim=0)               # rows = camera axes
        w2c = torch.eye(4, dtype=torch.float64)
        w2c[:3, :3] = Rwc
        w2c[:3, 3] = -Rwc @ c
        cams.append(SynthCamera(v, w2c.float(), fovx, fovy, width, height))
    return cams


@dataclass
class SynthScene:
    xyz: torch.Tensor            # [N,3]
    scaling: torch.Tensor        # [N,3]  log-scales
    rotation: torch.Tensor       # [N,4]  un-normalised (w,x,y,z)
    opacity: torch.Tensor        # [N,1]  logits
    features_dc: torch.Tensor    # [N,1,3]
    features_rest: torch.Tensor  # [N,15,3]
    semantic_feature: torch.Tensor  # [N,D]
    cameras: List[SynthCamera]


def make_scene(n: int, height: int, width: int, d: int, seed: int = 1234, n_views: int = 64,
               z_range=(2.0, 20.0), sigma_px_median: float = 2.0, with_sh: bool = True,
               feature_device=None) -> SynthScene:
    """Appendix B.  Means: pick a view, a uniform pixel and a depth U[z_range], un-project
    (=> uniform inside the union of the view frusta).  Scales: z*sigma_px/fx*exp(N(0,.4)) with
    sigma_px ~ LogNormal(ln 2, 0.7) => median projected sigma ~ 2 px."""
    g = torch.Generator().manual_seed(seed)
    cams = ring_cameras(n_views, width, height)
    fx = width / (2.0 * math.tan(cams[0].FoVx / 2.0))
    fy = height / (2.0 * math.tan(cams[0].FoVy / 2.0))
    view = torch.randint(0, n_views, (n,), generator=g)
    u = torch.rand(n, generator=g) * width
    v = torch.rand(n, generator=g) * height
    z = z_range[0] + (z_range[1] - z_range[0]) * torch.rand(n, generator=g)
    pc = torch.stack([(u - width / 2.0) * z / fx, (v - height / 2.0) * z / fy, z], dim=-1)
    c2w = torch.stack([torch.linalg.inv(c.world_view_transform.T.double()).float() for c in cams])
    Rcw = c2w[view, :3, :3]
    tcw = c2w[view, :3, 3]
    xyz = torch.einsum("nij,nj->ni", Rcw, pc) + tcw
    sigma_px = torch.exp(math.log(sigma_px_median) + 0.7 * torch.randn(n, generator=g))
    s = (z * sigma_px / fx)[:, None] * torch.exp(0.4 * torch.randn(n, 3, generator=g))
    scaling = torch.log(s)
    rotation = torch.randn(n, 4, generator=g)
    opacity = 1.5 * torch.randn(n, 1, generator=g)
    if feature_device is not None:
        # the [N, D] feature table drawn on the device (config 5: 10 GB — a minute of host RNG per
        # rank otherwise); geometry, cameras and everything else stay on the seeded host generator
        gd = torch.Generator(device=feature_device).manual_seed(seed + 7)
        sem = 0.1 * torch.randn(n, d, generator=gd, device=feature_device)
    else:
        sem = 0.1 * torch.randn(n, d, generator=g)
    if with_sh:
        f_dc = torch.rand(n, 1, 3, generator=g) * 2.0 - 1.0
        f_rest = 0.05 * torch.randn(n, 15, 3, generator=g)
    else:
        f_dc = torch.zeros(n, 1, 3)
        f_rest = torch.zeros(n, 15, 3)
    return SynthScene(xyz.contiguous(), scaling.contiguous(), rotation.contiguous(),
                      opacity.contiguous(), f_dc, f_rest, sem.contiguous(), cams)


def make_target(height: int, width: int, d: int, seed: int) -> torch.Tensor:
    """Distillation target N(0, 0.1^2), channel-last [H,W,D]."""
    g = torch.Generator().manual_seed(seed)
    return 0.1 * torch.randn(height, width, d, generator=g)


def config_scene(config_id: int, n_views: int = 64, with_sh: bool = False,
                 feature_device=None) -> SynthScene:
    """`feature_device`: draw the feature table there when it is huge (> 1.5e9 values, i.e. config 5);
    smaller tables always come from the host generator so that every run sees the same scene."""
    n, h, w, d = CONFIGS[config_id]
    dev = feature_device if n * d > 1_500_000_000 else None
    return make_scene(n, h, w, d, seed=1234 + config_id, n_views=n_views, with_sh=with_sh,
                      feature_device=dev)
