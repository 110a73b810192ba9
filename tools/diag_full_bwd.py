#!/usr/bin/env python
"""Error profile of the full blend backward vs the fp64 oracle (what test_full_backward_matches_oracle
asserts), per gradient and tolerance.  Run on the GPU box."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tests.helpers import front_scene, frac_bad, rel_err
from tests.test_gpu_parity import _stages, _blend_ref
from oracle import gags_oracle as O
from gags_b200 import rasterization as R

for D in (3, 4, 16, 64):
    W, H = 64, 48
    sc = front_scene(400, W, H, D, seed=200 + D)
    st = _stages(sc)
    g = torch.Generator().manual_seed(3)
    v_out = torch.randn(H, W, D, generator=g)
    v_alpha = torch.randn(H, W, generator=g)
    bg = torch.rand(D, generator=g)
    m2d, con, op, offs, ids = _blend_ref(st, None, None, W, H)
    leaves = [t.clone().requires_grad_(True) for t in (m2d, con, op, sc["colors"].double())]
    ref, ref_a, _ = O.blend_fwd(leaves[0], leaves[1], leaves[3], leaves[2], bg.double(), W, H, offs, ids)
    ((ref * v_out.double()).sum() + (ref_a * v_alpha.double()).sum()).backward()
    gl = [st["means2d"].clone().requires_grad_(True), st["conics"].clone().requires_grad_(True),
          st["opac"].clone().requires_grad_(True), sc["colors"].cuda().requires_grad_(True)]
    out, alphas, _ = R._Blend.apply(gl[0], gl[1], gl[2], gl[3], bg.cuda(), st["geom"], st["offsets"],
                                    st["flatten_ids"], W, H)
    ((out * v_out.cuda()).sum() + (alphas * v_alpha.cuda()).sum()).backward()
    for name, a, b in zip(("means2d", "conics", "opac", "colors"), gl, leaves):
        print(f"D={D:3d} {name:8s} rel_err {rel_err(a.grad, b.grad):.3e}  frac_bad@1e-4 {frac_bad(a.grad, b.grad, 1e-4):.2e} "
              f"@2e-4 {frac_bad(a.grad, b.grad, 2e-4):.2e} @5e-4 {frac_bad(a.grad, b.grad, 5e-4):.2e} @1e-3 {frac_bad(a.grad, b.grad, 1e-3):.2e}  n={a.grad.numel()}")
    print(f"D={D:3d} fwd rel_err {rel_err(out, ref):.3e} alpha {rel_err(alphas, ref_a):.3e}")
