"""SH helpers named as in /root/reference/utils/sh_utils.py: RGB <-> SH-DC (:114-118) and `eval_sh`
(:57-112).  The per-view SH colours of render() are evaluated by the CUDA kernel (csrc/sh.cu) inside
the rasteriser; `eval_sh` below is the general-purpose helper with the reference's signature, as a
basis-matrix contraction in PyTorch (any device, differentiable), for callers outside the hot path
(e.g. colour precomputation).  Same real-SH basis and sign convention; checked against the
reference's own function by tests/golden/reference_utils.npz."""
import torch

C0 = 0.28209479177387814
_C1 = 0.4886025119029199
_C2 = (1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792,
       0.5462742152960396)
_C3 = (-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
       -0.4570457994644658, 1.445305721320277, -0.5900435899266435)


def RGB2SH(rgb):
    return (rgb - 0.5) / C0


def SH2RGB(sh):
    return sh * C0 + 0.5


def sh_basis(deg: int, dirs: torch.Tensor) -> torch.Tensor:
    """[..., (deg+1)^2] real SH basis (degrees 0..3) at unit directions `dirs` [..., 3]."""
    if not 0 <= deg <= 3:
        raise ValueError("SH degree must be in 0..3")
    x, y, z = dirs.unbind(-1)
    cols = [torch.full_like(x, C0)]
    if deg >= 1:
        cols += [-_C1 * y, _C1 * z, -_C1 * x]
    if deg >= 2:
        xx, yy, zz = x * x, y * y, z * z
        cols += [_C2[0] * x * y, _C2[1] * y * z, _C2[2] * (2.0 * zz - xx - yy), _C2[3] * x * z,
                 _C2[4] * (xx - yy)]
    if deg >= 3:
        cols += [_C3[0] * y * (3 * xx - yy), _C3[1] * x * y * z, _C3[2] * y * (4 * zz - xx - yy),
                 _C3[3] * z * (2 * zz - 3 * xx - 3 * yy), _C3[4] * x * (4 * zz - xx - yy),
                 _C3[5] * z * (xx - yy), _C3[6] * x * (xx - 3 * yy)]
    return torch.stack(cols, dim=-1)


def eval_sh(deg: int, sh: torch.Tensor, dirs: torch.Tensor) -> torch.Tensor:
    """sh [..., C, K >= (deg+1)^2], dirs [..., 3] (unit length) -> [..., C]; no +0.5 offset, like the
    reference's eval_sh (the renderer adds it and clamps, gaussian_renderer / gsplat)."""
    nb = (deg + 1) ** 2
    if sh.shape[-1] < nb:
        raise ValueError("not enough SH coefficients for this degree")
    return (sh[..., :nb] * sh_basis(deg, dirs)[..., None, :]).sum(dim=-1)
