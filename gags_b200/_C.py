"""ctypes binding of include/gags_b200.h.  There is no fallback: if the shared library is missing
the import of this module raises, and every op that needs a GPU raises when CUDA is absent."""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GAGS_B200_LIB") or os.path.join(_HERE, "csrc", "libgags_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build it with `python -m gags_b200.build` "
        "(gags_b200 has no CPU or PyTorch fallback path)")

lib = C.CDLL(LIB_PATH)

GAGS_F_LOG_SCALES = 1
GAGS_F_LOGIT_OPACITY = 2


class Camera(C.Structure):
    _fields_ = [("viewmat", C.c_float * 16), ("viewmat_dev", C.c_void_p), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("width", C.c_int32), ("height", C.c_int32),
                ("eps2d", C.c_float), ("near_plane", C.c_float), ("far_plane", C.c_float),
                ("radius_clip", C.c_float), ("scaling_modifier", C.c_float), ("flags", C.c_int32)]


_p = C.c_void_p
_i32 = C.c_int32
_i64 = C.c_int64
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol include/gags_b200.h declares
SIGNATURES = {
    "gags_version": (C.c_char_p, []),
    "gags_build_arch": (C.c_char_p, []),
    "gags_error_string": (C.c_char_p, [C.c_int]),
    "gags_set_blend_impl": (C.c_int, [_i32]),
    "gags_get_blend_impl": (C.c_int, []),
    "gags_project_fwd": (C.c_int, [_p, _p, _p, _p, _i64, C.POINTER(Camera), _i32, _i32, _p, _p, _p,
                                   _p, _p, _p, _p, _p]),
    "gags_project_bwd": (C.c_int, [_p, _p, _p, _i64, C.POINTER(Camera), _p, _p, _p, _p, _p, _p, _p,
                                   _p, _p]),
    "gags_opacity_bwd": (C.c_int, [_p, _p, _i64, _p, _p]),
    "gags_sh_fwd": (C.c_int, [_i32, _p, _p, _p, _i32, _p, _i64, _p, _i32, _p]),
    "gags_sh_bwd": (C.c_int, [_i32, _p, _p, _p, _i32, _p, _i64, _p, _i32, _p, _p, _p]),
    "gags_tile_count": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p]),
    "gags_tile_scan_workspace_bytes": (_sz, [_i64]),
    "gags_tile_scan": (C.c_int, [_p, _i64, _p, _p, _p, _sz, _p]),
    "gags_tile_emit": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _p, _p, _p]),
    "gags_tile_bucket_max": (_i32, []),
    "gags_tile_bucket_count": (C.c_int, [_p, _p, _p, _i64, _i32, _i32, _p, _p, _p, _p, _p]),
    "gags_tile_bucket_sort": (C.c_int, [_p, _i32, _i32, _p, _i32, _p, _p, _p]),
    "gags_tile_bucket_sort_guarded": (C.c_int, [_p, _i32, _i32, _p, _i64, _p, _p, _p]),
    "gags_sort_pairs_workspace_bytes": (_sz, [_i64]),
    "gags_sort_pairs": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _p, _sz, C.POINTER(_i32), _p]),
    "gags_tile_offsets": (C.c_int, [_p, _i64, _i32, _p, _p]),
    "gags_blend_fwd": (C.c_int, [_p, _p, _i32, _p, _i32, _i32, _p, _p, _p, _p, _p, _p]),
    "gags_blend_bwd_features": (C.c_int, [_p, _i32, _i32, _i32, _p, _p, _p, _p, _p]),
    "gags_blend_cache_supported": (C.c_int, [_i32]),
    "gags_blend_last_ids_optional": (C.c_int, [_i32]),
    "gags_blend_cache_slots": (_i64, [_i64, _i32]),
    "gags_blend_fwd_cached": (C.c_int, [_p, _p, _i32, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                                        _p, _p]),
    "gags_blend_fwd_weights": (C.c_int, [_p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gags_blend_fwd_from_cache": (C.c_int, [_p, _i32, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gags_blend_bwd_features_cached": (C.c_int, [_i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gags_blend_bwd_features_cached_l1": (C.c_int, [_i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                                                    _p, _i32, _f, _p, _p, _p]),
    "gags_blend_bwd_full": (C.c_int, [_p, _p, _i32, _p, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                                      _p, _p, _p]),
    "gags_l1_loss_fused": (C.c_int, [_p, _p, _p, _i64, _i32, _f, _p, _p, _p]),
    "gags_l1_loss_segmap": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _f, _p, _p, _p]),
    "gags_l1_loss_sam": (C.c_int, [_p, _p, _p, _p, _i64, _i32, _i32, _f, _p, _p, _p, _p]),
    "gags_blend_bwd_features_cached_sam": (C.c_int, [_i32, _i32, _i32, _p, _p, _p, _p, _p, _p, _p, _p,
                                                     _p, _i32, _f, _p, _p, _p, _p]),
    "gags_pixel_loss_fwd": (C.c_int, [_i32, _p, _p, _i64, _i32, _p, _p, _p]),
    "gags_pixel_loss_bwd": (C.c_int, [_i32, _p, _p, _p, _p, _p, _i64, _i32, _p, _p]),
    "gags_scale_inplace": (C.c_int, [_p, _p, _i64, _p]),
    "gags_memset_zero": (C.c_int, [_p, _sz, _p]),
    "gags_zero_fill": (C.c_int, [_p, _i64, _p]),
    "gags_adam_step_multicast": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i64, C.c_double, C.c_double,
                                           C.c_double, C.c_double, _i32, _p]),
    "gags_set_peer_grid": (C.c_int, [_i32]),
    "gags_set_fwd_variant": (C.c_int, [_i32]),
    "gags_set_blend_pass": (C.c_int, [_i32]),
    "gags_set_bwd_sign_operand": (C.c_int, [_i32]),
    "gags_set_peer_unroll": (C.c_int, [_i32]),
    "gags_adam_step_peer": (C.c_int, [_i32, _i32, _p, _p, _p, _p, _i64, _i64, C.c_double,
                                      C.c_double, C.c_double, C.c_double, _i32, _p]),
    "gags_grad_allreduce_rows": (C.c_int, [_i32, _i32, _p, _p, _p, _p, _p, _i64, _i32, _p]),
    "gags_adam_step_consts": (C.c_int, [C.c_double, C.c_double, C.c_double, _i32, _p]),
    "gags_adam_lazy_rows": (C.c_int, [_p, _p, _p, _p, _p, _p, _p, _i64, _i32, _i32, _i32, C.c_double,
                                      C.c_double, C.c_double, _i32, _p]),
    "gags_adam_step_rows": (C.c_int, [_p, _p, _p, _p, _p, _i64, _i32, C.c_double, C.c_double,
                                      C.c_double, C.c_double, _i32, _p]),
    "gags_blend_cache_mark_rows": (C.c_int, [_i32, _i32, _p, _p, _p, _p, _p, _p]),
    "gags_adam_step": (C.c_int, [_p, _p, _p, _p, _i64, C.c_double, C.c_double, C.c_double,
                                 C.c_double, _i32, _i32, _p]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args


for _env, _fn in (("GAGS_B200_PEER_GRID", "gags_set_peer_grid"),
                  ("GAGS_B200_PEER_UNROLL", "gags_set_peer_unroll")):
    if os.environ.get(_env) and getattr(lib, _fn)(int(os.environ[_env])) != 0:
        raise ValueError(f"bad {_env}")
if os.environ.get("GAGS_B200_BLEND_IMPL"):           # 0 auto, 1 SIMT only, 2 tensor-core required
    if lib.gags_set_blend_impl(int(os.environ["GAGS_B200_BLEND_IMPL"])) != 0:
        raise ValueError("GAGS_B200_BLEND_IMPL must be 0, 1 or 2")
if os.environ.get("GAGS_B200_BWD_SIGN"):             # 1 exact sign operand (default), 0 hi / lo split
    if lib.gags_set_bwd_sign_operand(int(os.environ["GAGS_B200_BWD_SIGN"])) != 0:
        raise ValueError("GAGS_B200_BWD_SIGN must be 0 or 1")
if os.environ.get("GAGS_B200_BLEND_PASS"):            # 1 persistent (default), 0 one CTA per half tile
    if lib.gags_set_blend_pass(int(os.environ["GAGS_B200_BLEND_PASS"])) != 0:
        raise ValueError("GAGS_B200_BLEND_PASS must be 0 or 1")
if os.environ.get("GAGS_B200_FWD_VARIANT"):          # tuning / A-B switch, see include/gags_b200.h
    if lib.gags_set_fwd_variant(int(os.environ["GAGS_B200_FWD_VARIANT"])) != 0:
        raise ValueError("GAGS_B200_FWD_VARIANT must be 2, 3, 12 or 13")


def check(rc: int, what: str = "gags") -> None:
    if rc != 0:
        msg = lib.gags_error_string(rc).decode()
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("gags_b200 ops run on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


_launches = 0


def count_launch(n: int = 1) -> None:
    global _launches
    _launches += n


def launches() -> int:
    return _launches
