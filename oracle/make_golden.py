"""Generate tests/golden/reference_utils.npz by running the REFERENCE's own Python helpers
(imported unmodified from /root/reference) on seeded inputs.  Run in the build container only —
/root/reference does not exist on the GPU box; the committed .npz is what the tests read.

Pins the oracle (oracle/gags_oracle.py) against every piece of hot-path maths that exists in the
reference tree:
  utils/sh_utils.py:57-112        eval_sh            -> oracle.sh_basis
  utils/general_utils.py:78-110   build_rotation, build_scaling_rotation -> oracle.quat_to_rotmat,
                                                                            oracle.covariance3d
  utils/graphics_utils.py:38-77   getWorld2View2, fov2focal, getProjectionMatrix
                                                     -> oracle.world2view, oracle.intrinsics_from_fov

general_utils hard-codes device="cuda" (:83,:102); this container has no GPU, so torch.zeros is
shimmed to drop the device argument while those two functions run.  Nothing else is altered.
"""
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden",
                   "reference_utils.npz")


def main():
    sys.path.insert(0, REF)
    from utils import general_utils, graphics_utils, sh_utils  # noqa: E402

    g = torch.Generator().manual_seed(20241217)
    out = {}

    # --- SH ------------------------------------------------------------------------------------
    n = 64
    dirs = torch.nn.functional.normalize(torch.randn(n, 3, generator=g, dtype=torch.float64))
    sh = torch.randn(n, 3, 25, generator=g, dtype=torch.float64)      # [..., C, (deg+1)^2]
    out["sh_dirs"] = dirs.numpy()
    out["sh_coeffs"] = sh.numpy()
    for deg in range(5):
        out[f"sh_out_deg{deg}"] = sh_utils.eval_sh(deg, sh, dirs).numpy()
    out["rgb2sh"] = sh_utils.RGB2SH(torch.linspace(0, 1, 11, dtype=torch.float64)).numpy()

    # --- rotations / covariance (device shim) --------------------------------------------------
    quats = torch.randn(n, 4, generator=g)
    scales = torch.exp(torch.randn(n, 3, generator=g))
    real_zeros = torch.zeros

    def cpu_zeros(*a, **k):
        k.pop("device", None)
        return real_zeros(*a, **k)

    general_utils.torch.zeros = cpu_zeros
    try:
        R = general_utils.build_rotation(quats)
        L = general_utils.build_scaling_rotation(scales, quats)
        sym = general_utils.strip_symmetric(L @ L.transpose(1, 2))
    finally:
        general_utils.torch.zeros = real_zeros
    out["quats"] = quats.numpy()
    out["scales"] = scales.numpy()
    out["rotmats"] = R.numpy()
    out["L"] = L.numpy()
    out["cov_sym6"] = sym.numpy()
    out["inv_sigmoid_in"] = np.linspace(0.05, 0.95, 7).astype(np.float32)
    out["inv_sigmoid_out"] = general_utils.inverse_sigmoid(
        torch.from_numpy(out["inv_sigmoid_in"])).numpy()

    # --- cameras -------------------------------------------------------------------------------
    m = 8
    A = torch.randn(m, 3, 3, generator=g, dtype=torch.float64)
    Q, _ = torch.linalg.qr(A)
    Q = Q * torch.sign(torch.linalg.det(Q))[:, None, None]
    T = torch.randn(m, 3, generator=g, dtype=torch.float64)
    trans = torch.randn(m, 3, generator=g, dtype=torch.float64)
    sc = 0.5 + torch.rand(m, generator=g, dtype=torch.float64)
    out["cam_R"] = Q.numpy()
    out["cam_T"] = T.numpy()
    out["cam_trans"] = trans.numpy()
    out["cam_scale"] = sc.numpy()
    out["w2v"] = np.stack([graphics_utils.getWorld2View2(Q[i].numpy(), T[i].numpy(),
                                                         trans[i].numpy(), float(sc[i]))
                           for i in range(m)])
    fovs = np.array([0.5, 0.9, 1.0471975512, 1.4], dtype=np.float64)
    out["fovs"] = fovs
    out["focal_1920"] = np.array([graphics_utils.fov2focal(f, 1920) for f in fovs])
    out["fov_back"] = np.array([graphics_utils.focal2fov(graphics_utils.fov2focal(f, 1080), 1080)
                                for f in fovs])
    out["proj_matrix"] = graphics_utils.getProjectionMatrix(0.01, 100.0, 1.0, 0.7).numpy()

    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    np.savez_compressed(OUT, **out)
    print("wrote", os.path.normpath(OUT), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
