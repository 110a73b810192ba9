"""Shared test helpers: tiny hand-built scenes and tolerance checks.

Tolerance convention (BASELINE.json north_star: "within 1e-4 rel on float32, bit-exact on tile/sort
indices").  Float stages are compared against the fp64 oracle FED THE SAME STAGE INPUTS, with the
error measured relative to the reference tensor's scale (max |ref|): a blended pixel or a summed
gradient is a sum of many terms, and elementwise-relative error is meaningless where the sum
cancels to ~0.  Three kinds of assert appear in the GPU tests, each with its reason:
  * `rel_err(a, b) < 1e-4` (or tighter)  — every element within 1e-4 of scale: isolated stages whose
    control flow cannot differ (SH, Adam, losses, projection floats, cached vs recomputed weights).
  * `bad_pixels(...) <= k` / `frac_bad(...) < f` at 1e-4 — stages with a THRESHOLD in them
    (alpha >= 1/255, T <= 1e-4, sign(render - target) in the L1 gradient, ceil() in the radius): two
    correct implementations that round differently flip a decision at isolated elements, which
    changes those elements by far more than 1e-4.  The budget k / f is the counted allowance for
    such flips (a handful of pixels per ~10^4); everything else must hold 1e-4.
  * `frac_bad(..., 1e-3) < f` — END-TO-END checks through render(): the GPU projects in fp32, the
    oracle in fp64, so the two blends see means2d / conics that already differ by ~1e-6 relative,
    which a sharp Gaussian turns into ~1e-4..1e-3 of pixel value near its edge.  These tests pin the
    plumbing (argument order, activations, background, modes); the 1e-4 bar is carried by the
    per-stage tests above them.
"""
import math

import torch

from oracle import gags_oracle as O


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| relative to the scale of the reference tensor b (1e-4 bar of north_star)."""
    a, b = a.double().cpu(), b.double().cpu()
    scale = max(float(b.abs().max()), 1e-12)
    return float((a - b).abs().max()) / scale


def frac_bad(a, b, rtol=1e-4):
    """fraction of elements whose error exceeds rtol * tensor scale (threshold-flip tolerant)."""
    a, b = a.double().cpu(), b.double().cpu()
    scale = max(float(b.abs().max()), 1e-12)
    return float(((a - b).abs() > rtol * scale).double().mean())


def bad_pixels(a, b, rtol=1e-4):
    """number of pixels ([..., D] rows) with any channel off by more than rtol * tensor scale.
    Two implementations that round alpha differently may disagree on whether a Gaussian passes the
    alpha >= 1/255 or T <= 1e-4 threshold at an isolated pixel; the tests bound how many."""
    a, b = a.double().cpu(), b.double().cpu()
    scale = max(float(b.abs().max()), 1e-12)
    bad = ((a - b).abs() > rtol * scale).reshape(-1, a.shape[-1]).any(dim=-1)
    return int(bad.sum())


def identity_cam(width, height, fov_deg=60.0):
    fovx = math.radians(fov_deg)
    fx = width / (2 * math.tan(fovx / 2))
    fovy = 2 * math.atan(height / (2 * fx))
    return fovx, fovy, torch.eye(4)


def front_scene(n, width, height, d, seed=0, z=(3.0, 9.0), sigma_px=(1.5, 6.0), dtype=torch.float32):
    """n Gaussians in front of an identity camera, all inside the frustum; activated params."""
    g = torch.Generator().manual_seed(seed)
    fovx, fovy, vm = identity_cam(width, height)
    fx = width / (2 * math.tan(fovx / 2))
    zz = z[0] + (z[1] - z[0]) * torch.rand(n, generator=g)
    u = torch.rand(n, generator=g) * width
    v = torch.rand(n, generator=g) * height
    means = torch.stack([(u - width / 2) * zz / fx, (v - height / 2) * zz / fx, zz], -1)
    spx = sigma_px[0] + (sigma_px[1] - sigma_px[0]) * torch.rand(n, generator=g)
    scales = (zz * spx / fx)[:, None] * torch.exp(0.3 * torch.randn(n, 3, generator=g))
    quats = torch.randn(n, 4, generator=g)
    opac = torch.rand(n, generator=g) * 0.9 + 0.05
    colors = torch.randn(n, d, generator=g)
    K = O.intrinsics_from_fov(fovx, fovy, width, height)
    return dict(means=means.to(dtype), quats=quats.to(dtype), scales=scales.to(dtype),
                opacities=opac.to(dtype), colors=colors.to(dtype), viewmat=vm.to(dtype),
                K=K.to(dtype), width=width, height=height, fovx=fovx, fovy=fovy)
