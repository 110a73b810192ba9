"""The oracle against the golden vectors produced by the reference's own helpers
(oracle/make_golden.py -> tests/golden/reference_utils.npz).  CPU only."""
import math

import numpy as np
import torch

from oracle import gags_oracle as O


def test_sh_basis_matches_reference_eval_sh(golden):
    dirs = torch.from_numpy(golden["sh_dirs"])
    sh = torch.from_numpy(golden["sh_coeffs"])            # [n, 3, 25] (reference layout)
    for deg in range(5):
        nb = (deg + 1) ** 2
        B = O.sh_basis(deg, dirs)                          # [n, nb]
        got = (B[:, None, :] * sh[:, :, :nb]).sum(-1)
        ref = torch.from_numpy(golden[f"sh_out_deg{deg}"])
        assert torch.allclose(got, ref, rtol=1e-12, atol=1e-12), deg


def test_sh_colors_layout_and_offset(golden):
    # GaussianModel.get_features is [N, K, 3] (coeff-major), colours = max(sh + 0.5, 0)
    dirs = torch.from_numpy(golden["sh_dirs"])
    sh = torch.from_numpy(golden["sh_coeffs"])
    means = dirs * 3.0                                     # camera at the origin -> dir = mean/|mean|
    vm = torch.eye(4, dtype=torch.float64)
    radii = torch.ones(dirs.shape[0], dtype=torch.int32)
    got = O.sh_colors(3, means, vm, sh.permute(0, 2, 1).contiguous(), radii)
    ref = torch.clamp_min(torch.from_numpy(golden["sh_out_deg3"]) + 0.5, 0)
    assert torch.allclose(got, ref, rtol=1e-10, atol=1e-12)


def test_rotation_and_covariance_match_reference(golden):
    q = torch.from_numpy(golden["quats"])
    s = torch.from_numpy(golden["scales"])
    R = O.quat_to_rotmat(q)
    assert torch.allclose(R, torch.from_numpy(golden["rotmats"]), atol=2e-6)
    cov = O.covariance3d(q, s)
    sym = torch.from_numpy(golden["cov_sym6"])
    got6 = torch.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2],
                        cov[:, 2, 2]], 1)
    assert torch.allclose(got6, sym, rtol=1e-5, atol=1e-5 * float(sym.abs().max()))


def test_world2view_and_focal_match_reference(golden):
    for i in range(golden["cam_R"].shape[0]):
        w2v = O.world2view(golden["cam_R"][i], golden["cam_T"][i], golden["cam_trans"][i],
                           float(golden["cam_scale"][i]))
        assert np.allclose(w2v.numpy(), golden["w2v"][i], atol=1e-5)
    for f, ref in zip(golden["fovs"], golden["focal_1920"]):
        K = O.intrinsics_from_fov(float(f), float(f), 1920, 1080, torch.float64)
        assert math.isclose(float(K[0, 0]), float(ref), rel_tol=1e-12)
        assert float(K[0, 2]) == 960.0 and float(K[1, 2]) == 540.0


def test_product_helpers_match_reference(golden):
    """The product package's host-side mirrors (gags_b200.utils) against the same vectors."""
    from gags_b200.utils import general_utils as G, graphics_utils as GU, sh_utils as S
    q = torch.from_numpy(golden["quats"])
    s = torch.from_numpy(golden["scales"])
    assert torch.allclose(G.build_rotation(q), torch.from_numpy(golden["rotmats"]), atol=2e-6)
    assert torch.allclose(G.build_scaling_rotation(s, q), torch.from_numpy(golden["L"]), atol=1e-5)
    x = torch.from_numpy(golden["inv_sigmoid_in"])
    assert torch.allclose(G.inverse_sigmoid(x), torch.from_numpy(golden["inv_sigmoid_out"]))
    for i in range(golden["cam_R"].shape[0]):
        w2v = GU.getWorld2View2(golden["cam_R"][i], golden["cam_T"][i], golden["cam_trans"][i],
                                float(golden["cam_scale"][i]))
        assert np.allclose(w2v, golden["w2v"][i], atol=1e-5)
    assert np.allclose(GU.getProjectionMatrix(0.01, 100.0, 1.0, 0.7).numpy(), golden["proj_matrix"],
                       atol=1e-6)
    for f, ref in zip(golden["fovs"], golden["fov_back"]):
        assert math.isclose(GU.focal2fov(GU.fov2focal(float(f), 1080), 1080), float(ref))
    assert np.allclose(S.RGB2SH(torch.linspace(0, 1, 11, dtype=torch.float64)).numpy(),
                       golden["rgb2sh"])
