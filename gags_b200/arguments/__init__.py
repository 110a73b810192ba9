"""Default hyper-parameters of /root/reference/arguments/__init__.py:47-96 as plain objects
(the argparse plumbing itself is out of scope)."""
from types import SimpleNamespace


def OptimizationParams(**overrides):
    p = SimpleNamespace(iterations=30_000, position_lr_init=0.00016, position_lr_final=0.0000016,
                        position_lr_delay_mult=0.01, position_lr_max_steps=30_000,
                        feature_lr=0.0025, opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001,
                        semantic_feature_lr=0.001, percent_dense=0.01, lambda_dssim=0.2,
                        densification_interval=100, opacity_reset_interval=3000,
                        densify_from_iter=500, densify_until_iter=15_000,
                        densify_grad_threshold=0.0002)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def PipelineParams(**overrides):
    p = SimpleNamespace(convert_SHs_python=False, compute_cov3D_python=False, debug=False)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p


def ModelParams(**overrides):
    p = SimpleNamespace(sh_degree=3, source_path="", model_path="", images="images", resolution=-1,
                        white_background=False, data_device="cuda", eval=False, speedup=False)
    for k, v in overrides.items():
        setattr(p, k, v)
    return p
