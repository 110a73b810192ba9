"""RGB <-> SH-DC helpers named as in /root/reference/utils/sh_utils.py:114-118.  The SH basis
itself runs in csrc/sh.cu; eval_sh here only forwards to it for CUDA tensors."""
C0 = 0.28209479177387814


def RGB2SH(rgb):
    return (rgb - 0.5) / C0


def SH2RGB(sh):
    return sh * C0 + 0.5
