"""The operator boundary beneath render(): a B200-native stand-in for ``gsplat.rasterization`` as
the reference calls it (/root/reference/gaussian_renderer/__init__.py:56-70), C = 1, packed=False.

Every stage is a hand-written sm_100a kernel reached through the C-ABI of include/gags_b200.h
(``gags_b200._C``).  PyTorch only owns the buffers, the stream and the autograd graph:

    _Project   K1 (+K2 bwd)   means/quats/scales/opacities -> radii, means2d, depths, conics, opac
    _SHColors  K3 (+bwd)      SH coefficients -> RGB
    binning    K4-K6          tile count (fused in K1) -> scan -> emit -> radix sort -> offsets
    _Blend     K7 (+K8 bwd)   one launch for any D; feature-only backward when geometry is frozen;
                              with a lazily updated feature table (optim.LazyRows) the forward runs
                              as weights pass -> flag + catch up the rows it reads -> blend pass

There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes
import math
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import torch

from . import _C

TILE = 16

# Multi-view gradient accumulation (view-parallel training renders k views per optimiser step):
# when the colours ARE a leaf parameter, the feature backward can reduce straight into its .grad
# instead of into a fresh zeroed [N, D] buffer that autograd then adds to .grad (a 2 GB fill + a
# 6 GB read-modify-write per extra view at config 3).  Same arithmetic as AccumulateGrad's `+=`;
# what it skips are tensor / AccumulateGrad hooks on that parameter, hence opt-in
# (parallel.allreduce_grads does not use hooks).
direct_grad_accumulation = False

# last_ids (index of the last contributor per pixel) is read by the full geometry backward only.
# When no geometry gradient is needed the wide forward skips tracking it (cheaper transmittance
# chain, see csrc/blend_tc_common.cuh tc3_chain8) and info["last_ids"] is None; set this to keep it.
want_last_ids = False

# A lazily-updated feature table (optim.LazyRows) is rendered in two passes (see lazy_owners);
# False forces the flush + single-pass route (A/B switch for tests and bench).
two_pass_forward = True

# Keep the forward's blend-weight tiles for the feature backward (training with frozen geometry).
# The parity tests switch it off to exercise the recomputing backward kernels as well.
weight_cache = True

# The cache is sized by a closed-form bound (2 x (n_isects / 32 + n_tiles) tiles of 16 KB, ~10 GB at
# N = 2M / 1080p), of which a view touches ~15 %.  One grow-only buffer set per device is reused by
# every view: a view's tiles are dead once its backward has run, and a second forward before that
# backward (two live graphs) gets a fresh set instead.
_cache_pool: Dict = {}


class _CacheLease:
    """Buffers of one forward; returns them to the pool when the autograd node dies."""

    def __init__(self, dev, slots: int, n_half: int):
        key = (dev.type, dev.index)
        bufs = _cache_pool.pop(key, None)
        if bufs is None or bufs[2].numel() < slots or bufs[3].numel() < n_half + 1:
            cap = int(slots * 1.25) + 1024
            bufs = (torch.empty(cap * 16384, dtype=torch.uint8, device=dev),
                    torch.empty(cap * 32, dtype=torch.int32, device=dev),
                    torch.empty(cap, dtype=torch.int32, device=dev),
                    torch.empty(n_half + 1, dtype=torch.int32, device=dev))
        self.key, self.bufs = key, bufs

    def __del__(self):
        try:
            if self.key not in _cache_pool:
                _cache_pool[self.key] = self.bufs
        except Exception:
            pass

# Optional per-stage timing hooks (used by bench.py for the roofline numbers): when `stage_events`
# is a list, every stage boundary appends (name, torch.cuda.Event) recorded on the current stream.
stage_events = None


def _mark(name: str) -> None:
    if stage_events is not None:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        stage_events.append((name, ev))


# ------------------------------------------------------------------------------------------------
# one-view lookahead: geometry stage on a side stream
# ------------------------------------------------------------------------------------------------
# With frozen geometry (the shipped training loop, scene/gaussian_model.py:201-206) projection,
# tile keying and the sort of a view depend on nothing the optimiser writes, so they run on a
# high-priority side stream and overlap the previous view's loss / backward.  All work is still
# done every view; only its position in time changes.  Measured on B200 (tools/trace_step.py): the
# gain is modest (~0.3 ms of the 1.0 ms stage) because the sort and the L2-reduction-bound backward
# slow each other down; starting the stage beside Adam instead needs a shared-memory carveout on
# the Adam kernel and then costs Adam as much as it saves.  Rules that keep it safe:
#   * only when none of means / quats / scales / opacities requires grad;
#   * the side stream waits for the whole main stream the first time it sees an input tensor
#     (pointer + version; a strong reference is kept so the address cannot be recycled), and
#     otherwise only for the end of the previous view's forward blend — which also bounds the
#     lookahead (and the host run-ahead, through the n_isects readback) to one view;
#   * every buffer produced on the side stream is record_stream()ed on the consuming stream.
import os as _os
# Off by default since the loss moved into the backward (fused_l1_backward): without the HBM-bound
# loss pass to hide behind, the stage only ever overlaps the L2-reduction-bound backward and Adam,
# and both lose more than the stage gains (measured: 6.06 ms/step off, 6.57 ms/step on).
lookahead = _os.environ.get("GAGS_B200_LOOKAHEAD", "0") != "0"
# Zero the backward's [N, D] accumulation buffer on a second stream, launched right AFTER the
# forward blend kernel: the forward is issue / shared-memory bound and leaves thread slots and HBM
# bandwidth free, so the fill's CTAs run in its shadow instead of in front of the backward.
prezero_overlap = _os.environ.get("GAGS_B200_PREZERO", "1") != "0"
_zero_streams: Dict = {}


_readback: Dict = {}


def _readback_state(dev):
    """(copy stream, pinned 2-int buffer) for the n_isects readback beside a running sort."""
    st = _readback.get(dev.index)
    if st is None:
        st = _readback[dev.index] = (torch.cuda.Stream(device=dev, priority=-1),
                                     torch.empty(2, dtype=torch.int32).pin_memory())
    return st


# Hand-offs between rasterize_view() and the autograd nodes it creates (the pre-zeroed gradient
# buffer, the context of the forward that kept its weight tiles).  They live for the duration of ONE
# rasterize_view() call on ONE thread: thread-local, and cleared when the call ends however it ends.
import threading as _threading
_tls = _threading.local()


def _early_prezero(colors, geo_in, sh_degree) -> None:
    """Start zeroing the backward's [N, D] accumulation buffer at the very beginning of the view,
    on its own stream: the projection / scatter / sort kernels in front of the blend are latency
    bound and leave HBM idle, whereas the forward and backward blends fill the register file and
    the optimiser pass is HBM bound — nothing can run in THEIR shadow (measured: a fill launched
    right behind the forward starts when the forward drains)."""
    _tls.pending_prezero = None
    if not (prezero_overlap and stage_events is None and weight_cache and sh_degree is None
            and torch.is_grad_enabled() and colors.is_cuda and colors.requires_grad
            and colors.dim() == 2 and not any(t.requires_grad for t in geo_in)):
        return
    N, D = colors.shape
    if not _C.lib.gags_blend_cache_supported(D):
        return
    if (_direct(colors) and colors.is_leaf and colors.grad is not None):
        return                                        # the backward will reduce straight into .grad
    dev = colors.device
    zs = _zero_stream(dev)
    # not earlier than the main stream gets here (the host runs a view ahead of the device and the
    # fill would otherwise land between the previous view's blends)
    ev0 = torch.cuda.Event()
    ev0.record(torch.cuda.current_stream(dev))
    zs.wait_event(ev0)
    with torch.cuda.stream(zs):
        vz = torch.empty(N, D, dtype=torch.float32, device=dev)
        _C.check(_C.lib.gags_zero_fill(_C.ptr(vz), vz.numel() * 4, zs.cuda_stream),
                 "gags_zero_fill")                    # small-grid fill: leaves the SMs' slots free
        _C.count_launch()
        evz = torch.cuda.Event()
        evz.record(zs)
    _tls.pending_prezero = (vz, evz, torch.cuda.current_stream(dev))


def _zero_stream(dev):
    s = _zero_streams.get(dev.index)
    if s is None:
        s = _zero_streams[dev.index] = torch.cuda.Stream(device=dev)
    return s
_side: Dict = {}


def _side_state(dev):
    st = _side.get(dev.index)
    if st is None:
        st = dict(stream=torch.cuda.Stream(device=dev, priority=-1), ev_prev=None,
                  seen=OrderedDict())
        _side[dev.index] = st
    return st


def register_static(*tensors) -> None:
    """Declare device tensors (camera matrices, frozen geometry) whose contents are final: after one
    device synchronisation the side stream may read them without waiting for the main stream.
    Optional — an unregistered tensor is simply waited for the first time it is seen."""
    devs = {t.device for t in tensors if t.is_cuda}
    for d in devs:
        torch.cuda.synchronize(d)
    for t in tensors:
        if t.is_cuda:
            _all_seen(_side_state(t.device), (t,))


def _all_seen(st, tensors) -> bool:
    seen, ok = st["seen"], True
    for t in tensors:
        key = (t.data_ptr(), t._version, t.numel())
        if key in seen:
            seen.move_to_end(key)
        else:
            ok = False
            seen[key] = t
    while len(seen) > 1024:
        seen.popitem(last=False)
    return ok


# ------------------------------------------------------------------------------------------------
# camera struct
# ------------------------------------------------------------------------------------------------
def make_camera(viewmat: torch.Tensor, fx: float, fy: float, cx: float, cy: float, width: int,
                height: int, eps2d: float = 0.3, near_plane: float = 0.01,
                far_plane: float = 1e10, radius_clip: float = 0.0, scaling_modifier: float = 1.0,
                flags: int = 0) -> Tuple[_C.Camera, torch.Tensor]:
    """`viewmat` [4,4] world->camera.  A CUDA view matrix is read by the kernels straight from
    device memory (no host round trip); a CPU one is copied into the struct."""
    cam = _C.Camera()
    keep = None
    if viewmat.is_cuda:
        keep = viewmat.detach().to(torch.float32).contiguous()
        cam.viewmat_dev = keep.data_ptr()
    else:
        vm = viewmat.detach().to(torch.float32).contiguous().reshape(-1).tolist()
        for i in range(16):
            cam.viewmat[i] = vm[i]
        cam.viewmat_dev = None
    cam.fx, cam.fy, cam.cx, cam.cy = float(fx), float(fy), float(cx), float(cy)
    cam.width, cam.height = int(width), int(height)
    cam.eps2d, cam.near_plane, cam.far_plane = float(eps2d), float(near_plane), float(far_plane)
    cam.radius_clip = float(radius_clip)
    cam.scaling_modifier = float(scaling_modifier)
    cam.flags = int(flags)
    return cam, keep


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise ValueError(f"expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


# ------------------------------------------------------------------------------------------------
# K1 / K2
# ------------------------------------------------------------------------------------------------
class _Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means, quats, scales, opacities, cam, cam_keep, tile_w, tile_h):
        _C.require_cuda(means, quats, scales, opacities)
        means, quats, scales = _f32c(means), _f32c(quats), _f32c(scales)
        opacities = _f32c(opacities).reshape(-1)
        N = means.shape[0]
        if means.shape != (N, 3) or quats.shape != (N, 4) or scales.shape != (N, 3) \
                or opacities.shape != (N,):
            raise ValueError("means [N,3], quats [N,4], scales [N,3], opacities [N] expected")
        dev = means.device
        radii = torch.empty(N, dtype=torch.int32, device=dev)
        means2d = torch.empty(N, 2, dtype=torch.float32, device=dev)
        depths = torch.empty(N, dtype=torch.float32, device=dev)
        conics = torch.empty(N, 3, dtype=torch.float32, device=dev)
        opac = torch.empty(N, dtype=torch.float32, device=dev)
        tiles = torch.empty(N, dtype=torch.int32, device=dev)
        geom = torch.empty(N, 8, dtype=torch.float32, device=dev)
        _C.check(_C.lib.gags_project_fwd(_C.ptr(means), _C.ptr(quats), _C.ptr(scales),
                                         _C.ptr(opacities), N, ctypes.byref(cam), tile_w, tile_h,
                                         _C.ptr(radii), _C.ptr(means2d), _C.ptr(depths),
                                         _C.ptr(conics), _C.ptr(opac), _C.ptr(tiles), _C.ptr(geom),
                                         _C.stream_ptr()), "gags_project_fwd")
        _C.count_launch()
        ctx.cam, ctx.cam_keep = cam, cam_keep
        ctx.save_for_backward(means, quats, scales, radii, conics, opac)
        ctx.mark_non_differentiable(radii, tiles, geom)
        return radii, means2d, depths, conics, opac, tiles, geom

    @staticmethod
    def backward(ctx, _vr, v_means2d, v_depths, v_conics, v_opac, _vt, _vg):
        means, quats, scales, radii, conics, opac = ctx.saved_tensors
        N = means.shape[0]
        dev = means.device
        need_geo = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        v_means = v_quats = v_scales = v_logit = None
        if need_geo:
            z2 = v_means2d.contiguous() if v_means2d is not None else torch.zeros(N, 2, device=dev)
            z3 = v_conics.contiguous() if v_conics is not None else torch.zeros(N, 3, device=dev)
            zd = v_depths.contiguous() if v_depths is not None else None
            v_means = torch.empty(N, 3, device=dev)
            v_quats = torch.empty(N, 4, device=dev)
            v_scales = torch.empty(N, 3, device=dev)
            _C.check(_C.lib.gags_project_bwd(_C.ptr(means), _C.ptr(quats), _C.ptr(scales), N,
                                             ctypes.byref(ctx.cam), _C.ptr(radii), _C.ptr(conics),
                                             _C.ptr(z2), _C.ptr(zd), _C.ptr(z3), _C.ptr(v_means),
                                             _C.ptr(v_quats), _C.ptr(v_scales), _C.stream_ptr()),
                     "gags_project_bwd")
            _C.count_launch()
        if ctx.needs_input_grad[3] and v_opac is not None:
            if ctx.cam.flags & _C.GAGS_F_LOGIT_OPACITY:
                v_logit = torch.empty(N, device=dev)
                _C.check(_C.lib.gags_opacity_bwd(_C.ptr(opac), _C.ptr(v_opac.contiguous()), N,
                                                 _C.ptr(v_logit), _C.stream_ptr()),
                         "gags_opacity_bwd")
                _C.count_launch()
            else:
                v_logit = v_opac
        return v_means, v_quats, v_scales, v_logit, None, None, None, None


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
class _SHColors(torch.autograd.Function):
    @staticmethod
    def forward(ctx, degree, means, campos, coeffs, radii):
        _C.require_cuda(means, coeffs, campos)
        means, coeffs = _f32c(means), _f32c(coeffs)
        N, K = coeffs.shape[0], coeffs.shape[1]
        if coeffs.shape[2] != 3 or K < (degree + 1) ** 2:
            raise ValueError("SH coefficients must be [N, K>=(deg+1)^2, 3]")
        colors = torch.empty(N, 3, dtype=torch.float32, device=means.device)
        campos = _f32c(campos)
        _C.check(_C.lib.gags_sh_fwd(degree, _C.ptr(means), _C.ptr(campos), _C.ptr(coeffs), K,
                                    _C.ptr(radii), N, _C.ptr(colors), 3, _C.stream_ptr()),
                 "gags_sh_fwd")
        _C.count_launch()
        ctx.degree = degree
        ctx.save_for_backward(means, campos, coeffs, radii)
        return colors

    @staticmethod
    def backward(ctx, v_colors):
        means, campos, coeffs, radii = ctx.saved_tensors
        N, K = coeffs.shape[0], coeffs.shape[1]
        v_coeffs = torch.empty_like(coeffs)
        v_means = torch.zeros_like(means) if ctx.needs_input_grad[1] else None
        _C.check(_C.lib.gags_sh_bwd(ctx.degree, _C.ptr(means), _C.ptr(campos), _C.ptr(coeffs), K,
                                    _C.ptr(radii), N, _C.ptr(v_colors.contiguous()), 3,
                                    _C.ptr(v_coeffs), _C.ptr(v_means), _C.stream_ptr()),
                 "gags_sh_bwd")
        _C.count_launch()
        return None, v_means, None, v_coeffs, None


# ------------------------------------------------------------------------------------------------
# K4-K6 binning (integer stage, no autograd)
# ------------------------------------------------------------------------------------------------
def tile_bits(n_tiles: int) -> int:
    return int(math.floor(math.log2(n_tiles))) + 1


# grow-only capacity of the intersection buffers, per device (index -> entries)
_isect_capacity: Dict[int, int] = {}
# True = K4-K6 through the tile-bucketed segmented sort (csrc/tile_buckets.cu); False = count /
# scan / emit / global cub radix sort / offsets (csrc/tiles.cu).  Both give identical arrays
# (tests/test_gpu_parity.py::test_tile_stage_is_bit_exact).  Measured at config 3 on B200: 0.57 ms
# bucketed (scatter 0.21 + scan 0.01 + block sort 0.35) vs 0.89 ms global (6 onesweep passes); a
# tile with more than gags_tile_bucket_max() intersections falls back to the global path.
bucket_sort = True
# enqueue the bucket sort before the n_isects readback once a capacity is known (see bin_and_sort)
speculative_sort = True


@torch.no_grad()
def bin_and_sort(means2d, radii, depths, tiles_touched, tile_w: int, tile_h: int,
                 after_count=None) -> Dict:
    """tiles_touched -> scan -> emit -> stable radix sort -> offsets.  One host sync (n_isects).
    `after_count()` (optional) runs on the host right after that readback, before the emit / sort
    kernels are enqueued (the lookahead uses it to order them behind the main stream)."""
    N = radii.shape[0]
    dev = radii.device
    st = _C.stream_ptr()
    cap_key = dev.index if dev.index is not None else torch.cuda.current_device()
    known_cap = _isect_capacity.get(cap_key, 0)
    if bucket_sort:
        # tile-bucketed segmented sort: scatter + scan give the offsets and n_isects directly
        n_tiles = tile_w * tile_h
        bmax = int(_C.lib.gags_tile_bucket_max())
        count = torch.empty(n_tiles, dtype=torch.int32, device=dev)
        bucket = torch.empty(n_tiles * bmax, dtype=torch.int64, device=dev)
        offsets = torch.empty(n_tiles + 1, dtype=torch.int32, device=dev)
        stats = torch.empty(2, dtype=torch.int32, device=dev)
        _C.check(_C.lib.gags_tile_bucket_count(_C.ptr(means2d), _C.ptr(radii), _C.ptr(depths), N,
                                               tile_w, tile_h, _C.ptr(count), _C.ptr(bucket),
                                               _C.ptr(offsets), _C.ptr(stats), st),
                 "gags_tile_bucket_count")
        _C.count_launch(2)
        # With a capacity known from earlier views the sort is enqueued BEFORE the counts are read
        # back, so the device sorts during the host round trip; its guard skips tiles that do not
        # fit, and the (rare) view that outgrew the capacity is sorted again below.
        keys = vals = None
        spec_cap = known_cap if (speculative_sort and after_count is None) else 0
        if spec_cap > 0:
            main = torch.cuda.current_stream(dev)
            ev_counts = torch.cuda.Event()
            ev_counts.record(main)                            # counts + scan done
            keys = torch.empty(spec_cap, dtype=torch.int64, device=dev)
            vals = torch.empty(spec_cap, dtype=torch.int32, device=dev)
            _C.check(_C.lib.gags_tile_bucket_sort_guarded(_C.ptr(bucket), tile_w, tile_h,
                                                          _C.ptr(offsets), spec_cap, _C.ptr(keys),
                                                          _C.ptr(vals), st),
                     "gags_tile_bucket_sort_guarded")
            _C.count_launch(2)                                # small- and large-bucket launch
            # the readback goes through a second stream that waits for the scan only, not the sort
            cs, pinned = _readback_state(dev)
            cs.wait_event(ev_counts)
            with torch.cuda.stream(cs):
                pinned.copy_(stats, non_blocking=True)
            stats.record_stream(cs)
            cs.synchronize()                                  # the one host sync of the pipeline
            n, max_bucket = (int(v) for v in pinned.tolist())
        else:
            n, max_bucket = (int(v) for v in stats.tolist())  # the one host sync of the pipeline
        if after_count is not None:
            after_count()
            after_count = None
        if max_bucket <= bmax:
            if spec_cap > 0 and n <= spec_cap:
                return dict(n_isects=n, isect_ids=keys[:n], flatten_ids=vals[:n], offsets=offsets,
                            cum_tiles=None, _bases=(keys, vals, offsets))
            if n > known_cap:
                known_cap = _isect_capacity[cap_key] = int(n * 1.2) + 1024
            cap = known_cap
            keys = torch.empty(cap, dtype=torch.int64, device=dev)
            vals = torch.empty(cap, dtype=torch.int32, device=dev)
            if n > 0:
                _C.check(_C.lib.gags_tile_bucket_sort(_C.ptr(bucket), tile_w, tile_h,
                                                      _C.ptr(offsets), max_bucket, _C.ptr(keys),
                                                      _C.ptr(vals), st), "gags_tile_bucket_sort")
                _C.count_launch(2)
            return dict(n_isects=n, isect_ids=keys[:n], flatten_ids=vals[:n], offsets=offsets,
                        cum_tiles=None, _bases=(keys, vals, offsets))
        # a tile with more intersections than a CTA sorts in shared memory: global radix sort
    cum = torch.empty(N, dtype=torch.int32, device=dev)
    n_dev = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = _C.lib.gags_tile_scan_workspace_bytes(N)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _C.check(_C.lib.gags_tile_scan(_C.ptr(tiles_touched), N, _C.ptr(cum), _C.ptr(n_dev),
                                   _C.ptr(ws), ws_bytes, st), "gags_tile_scan")
    _C.count_launch(2)
    n = int(n_dev.item())                     # the one host sync of the pipeline
    if after_count is not None:
        after_count()
    n_tiles = tile_w * tile_h
    # n varies by a few percent from view to view; allocating a grow-only capacity instead keeps
    # every view's request the same size, so the caching allocator recycles blocks instead of
    # calling cudaMalloc (a device-wide sync) every other view
    if n > known_cap:
        known_cap = _isect_capacity[cap_key] = int(n * 1.2) + 1024
    cap = known_cap
    keys_a = torch.empty(cap, dtype=torch.int64, device=dev)
    keys_b = torch.empty(cap, dtype=torch.int64, device=dev)
    vals_a = torch.empty(cap, dtype=torch.int32, device=dev)
    vals_b = torch.empty(cap, dtype=torch.int32, device=dev)
    offsets = torch.empty(n_tiles + 1, dtype=torch.int32, device=dev)
    if n > 0:
        _C.check(_C.lib.gags_tile_emit(_C.ptr(means2d), _C.ptr(radii), _C.ptr(depths), _C.ptr(cum),
                                       N, tile_w, tile_h, _C.ptr(keys_a), _C.ptr(vals_a), st),
                 "gags_tile_emit")
        sws_bytes = _C.lib.gags_sort_pairs_workspace_bytes(cap)
        sws = torch.empty(sws_bytes, dtype=torch.uint8, device=dev)
        sel = ctypes.c_int32(0)
        _C.check(_C.lib.gags_sort_pairs(_C.ptr(keys_a), _C.ptr(keys_b), _C.ptr(vals_a),
                                        _C.ptr(vals_b), n, 32 + tile_bits(n_tiles), _C.ptr(sws),
                                        sws_bytes, ctypes.byref(sel), st), "gags_sort_pairs")
        _C.count_launch(8)
        if sel.value == 1:
            keys_a, vals_a = keys_b, vals_b
    _C.check(_C.lib.gags_tile_offsets(_C.ptr(keys_a), n, n_tiles, _C.ptr(offsets), st),
             "gags_tile_offsets")
    _C.count_launch()
    return dict(n_isects=n, isect_ids=keys_a[:n], flatten_ids=vals_a[:n], offsets=offsets,
                cum_tiles=cum, _bases=(keys_a, vals_a, offsets, cum))


# ------------------------------------------------------------------------------------------------
# K7 / K8
# ------------------------------------------------------------------------------------------------
# data_ptr of a persistent `.grad` buffer -> event after which it is zeroed (parallel.PeerAdam)
sink_ready_events: Dict = {}
# data_ptr of a feature table -> event after which its latest optimiser update has landed (PeerAdam
# runs the multi-GPU exchange on its own stream, beside the next view's projection / sort)
param_ready_events: Dict = {}


class RowFlags:
    """Per-row "this gradient row may be non-zero" flags of a persistent `.grad` buffer
    (optim.FusedAdam(sparse_rows=True), parallel.SparsePeerAdam).  `flags` uint8 [N]: 0 promises an
    all-zero row.  `dirty`: a backward has written into the buffer since the owner last consumed it."""
    __slots__ = ("flags", "dirty")

    def __init__(self, flags):
        self.flags, self.dirty = flags, False


# data_ptr of a persistent `.grad` buffer -> RowFlags.  The owner (the optimiser) keeps the buffer
# alive for as long as the entry exists, so the address cannot be recycled under it.
row_flags: Dict[int, RowFlags] = {}


# data_ptr of a feature table -> optim.LazyRows: the table is updated lazily (rows no view touched
# are behind by some optimiser steps).  A forward that keeps its weight tiles runs in two passes —
# weights pass, flag + catch up exactly the rows it blends, blend pass; any other forward flushes the
# whole table first.
lazy_owners: Dict = {}


def _direct(colors) -> bool:
    """Reduce this tensor's feature gradient straight into its `.grad`?  Process-wide switch or the
    per-parameter opt-in an optimiser sets on the parameters it keeps a persistent gradient for."""
    return direct_grad_accumulation or getattr(colors, "_gags_direct_grad", False)


def _mark_rows(v_colors, cache, offsets, width: int, height: int) -> None:
    """After a feature backward that accumulated into `v_colors`: if that buffer carries row flags,
    flag the rows this view can have touched (cached backward: the rows of the batches the forward
    kept; any other kernel: all rows)."""
    rf = row_flags.get(v_colors.data_ptr()) if (row_flags and v_colors is not None) else None
    if rf is None:
        return
    rf.dirty = True
    if cache is not None:
        _C.check(_C.lib.gags_blend_cache_mark_rows(width, height, _C.ptr(offsets), _C.ptr(cache[1]),
                                                   _C.ptr(cache[2]), _C.ptr(cache[3]),
                                                   _C.ptr(rf.flags), _C.stream_ptr()),
                 "gags_blend_cache_mark_rows")
        _C.count_launch()
    else:
        rf.flags.fill_(1)


def _take_grad_buffer(ctx, need_col: bool, need_geo: bool, N: int, D: int, dev):
    """The [N, D] buffer the feature backward accumulates into, and the leaf it belongs to when the
    reduction goes straight into `.grad` (direct_grad_accumulation)."""
    sink = ctx.sink if (need_col and not need_geo) else None
    if sink is not None and sink.grad is not None and not (
            sink.grad.is_contiguous() and sink.grad.dtype == torch.float32):
        sink = None                              # cannot reduce in place: autograd adds the result
    if sink is not None and sink.grad is not None:
        v_colors = sink.grad                     # accumulate in place: nothing to zero or add
        ctx.prezero = None
        ev = sink_ready_events.pop(v_colors.data_ptr(), None)
        if ev is not None:                       # a persistent buffer re-zeroed on another stream
            torch.cuda.current_stream(dev).wait_event(ev)
    elif need_col and ctx.prezero is not None:
        v_colors, evz, _ = ctx.prezero
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(evz)
        v_colors.record_stream(cur)
        ctx.prezero = None
    else:
        v_colors = torch.zeros(N, D, device=dev) if need_col else None
    return v_colors, sink


class _FusedHandle:
    """What fused_l1_backward needs from one cached forward (attached to the render tensor)."""
    __slots__ = ("ctx", "cols", "offsets", "render_ptr")

    def __init__(self, ctx, cols, offsets, render_ptr):
        self.ctx, self.cols, self.offsets, self.render_ptr = ctx, cols, offsets, render_ptr


def _fused_handle(render_dhw):
    """(handle, channel-last render) when `render_dhw` is the output of a forward that kept its
    weight tiles and has not been consumed yet, else (None, None)."""
    h = getattr(render_dhw, "_gags_fused", None)
    r = render_dhw.permute(1, 2, 0) if render_dhw.dim() == 3 else None
    if (h is not None and h.ctx.lease is not None and r is not None and r.is_contiguous()
            and r.data_ptr() == h.render_ptr):
        return h, r
    return None, None


def _finish_fused_backward(render_dhw, h, v_colors, sink) -> None:
    """Hand the feature gradient of a fused loss + backward call to its owner."""
    ctx = h.ctx
    ctx.lease = None               # this view's weight tiles are spent: back to the pool
    render_dhw._gags_fused = None
    if sink is not None:
        if sink.grad is None:
            sink.grad = v_colors
    elif h.cols.requires_grad:
        cols = h.cols
        if cols.is_leaf:
            # what AccumulateGrad does with a gradient nobody else holds: adopt it (no 2 GB copy)
            if cols.grad is None:
                cols.grad = v_colors
            else:
                cols.grad.add_(v_colors)
        else:
            torch.autograd.backward([cols], [v_colors])  # through whatever produced the features


def fused_l1_backward(render_dhw, seg_hw, emb, mask_hw=None):
    """loss = mean(|render - emb[seg]| * mask) AND its backward into the feature table in one
    kernel: equivalent to `loss = l1_loss_segmap_fused(render, seg, emb, mask); loss.backward()`
    (train.py:162-165 with the compact target), but the loss gradient is formed inside the cached
    feature backward's staging warps, so the [H,W,D] gradient map (2 GB at config 3) is neither
    written nor read.  `render_dhw` must be the tensor render() / rasterize_view() returned for a
    forward that kept its weight tiles (frozen geometry, trainable features); anything else falls
    back to the two-kernel form.  Returns the detached loss."""
    from .utils.loss_utils import _mask_f32, l1_loss_segmap_fused
    h, r = _fused_handle(render_dhw)
    if h is None or seg_hw.dtype != torch.int32 or emb.dtype != torch.float32 or emb.dim() != 2:
        loss = l1_loss_segmap_fused(render_dhw, seg_hw, emb, mask_hw)
        loss.backward()
        return loss.detach()
    ctx = h.ctx
    width, height, D, N = ctx.dims
    if seg_hw.shape != (height, width) or emb.shape[1] != D:
        raise ValueError("seg must be int32 [H,W] and emb float32 [n_seg, D]")
    _C.require_cuda(seg_hw, emb, mask_hw)
    dev = r.device
    sg, em = seg_hw.contiguous(), emb.contiguous()
    m = _mask_f32(mask_hw, height, width)
    _mark("bwd_start")
    v_colors, sink = _take_grad_buffer(ctx, True, False, N, D, dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    _mark("bwd_zero")
    cache = ctx.lease.bufs
    numel = float(height * width * D)
    _C.check(_C.lib.gags_blend_bwd_features_cached_l1(
        D, width, height, _C.ptr(h.offsets), _C.ptr(cache[0]), _C.ptr(cache[1]), _C.ptr(cache[2]),
        _C.ptr(cache[3]), _C.ptr(r), _C.ptr(sg), _C.ptr(em), _C.ptr(m), em.shape[0], 1.0 / numel,
        _C.ptr(loss), _C.ptr(v_colors), _C.stream_ptr()), "gags_blend_bwd_features_cached_l1")
    _C.count_launch((D + 255) // 256)
    if not ctx.rows_marked:
        _mark_rows(v_colors, cache, h.offsets, width, height)
    _mark("blend_bwd")
    _finish_fused_backward(render_dhw, h, v_colors, sink)
    return loss[0] / numel


def fused_sam_backward(render_dhw, seg3, emb, scale_map, want_scale_grad: bool = False):
    """The same single-kernel loss + backward against the reference's FULL distillation target
    (read_sam_clip_feature + l1_loss, /root/reference/scene/dataset_readers.py:54-121 and
    /root/reference/train.py:162-163): `seg3` int32 [3,H,W] (SAM levels s, m, l; -1 = none), `emb`
    [n_seg, D], `scale_map` float32 [3,H,W] (the scale decoder's output).  Returns (loss,
    d loss / d scale_map or None).  Falls back to loss_utils.l1_loss_sam_fused + backward when the
    render did not keep its weight tiles."""
    from .utils.loss_utils import l1_loss_sam_fused
    h, r = _fused_handle(render_dhw)
    ok = (h is not None and seg3.dtype == torch.int32 and seg3.dim() == 3 and seg3.shape[0] == 3
          and emb.dtype == torch.float32 and emb.dim() == 2
          and scale_map.dtype == torch.float32 and scale_map.shape == seg3.shape)
    if ok:
        width, height, D, N = h.ctx.dims
        ok = tuple(seg3.shape[1:]) == (height, width) and emb.shape[1] == D \
            and not (want_scale_grad and D % 128 != 0)
    if not ok:
        sm = scale_map.detach().requires_grad_(True) if want_scale_grad else scale_map
        loss = l1_loss_sam_fused(render_dhw, seg3, emb, sm)
        loss.backward()
        return loss.detach(), (sm.grad if want_scale_grad else None)
    ctx = h.ctx
    _C.require_cuda(seg3, emb, scale_map)
    dev = r.device
    sg, em, sc = seg3.contiguous(), emb.contiguous(), scale_map.detach().contiguous()
    v_scale = torch.zeros_like(sc) if want_scale_grad else None
    _mark("bwd_start")
    v_colors, sink = _take_grad_buffer(ctx, True, False, N, D, dev)
    loss = torch.zeros(1, dtype=torch.float32, device=dev)
    _mark("bwd_zero")
    cache = ctx.lease.bufs
    numel = float(height * width * D)
    _C.check(_C.lib.gags_blend_bwd_features_cached_sam(
        D, width, height, _C.ptr(h.offsets), _C.ptr(cache[0]), _C.ptr(cache[1]), _C.ptr(cache[2]),
        _C.ptr(cache[3]), _C.ptr(r), _C.ptr(sg), _C.ptr(em), _C.ptr(sc), em.shape[0], 1.0 / numel,
        _C.ptr(loss), _C.ptr(v_scale), _C.ptr(v_colors), _C.stream_ptr()),
        "gags_blend_bwd_features_cached_sam")
    _C.count_launch((D + 255) // 256)
    if not ctx.rows_marked:
        _mark_rows(v_colors, cache, h.offsets, width, height)
    _mark("blend_bwd")
    _finish_fused_backward(render_dhw, h, v_colors, sink)
    return loss[0] / numel, v_scale


class _Blend(torch.autograd.Function):
    """means2d / conics / opac are the differentiable handles of the projection outputs; the
    kernels read the same values from the packed `geom` record."""

    @staticmethod
    def forward(ctx, means2d, conics, opac, colors, background, geom, offsets, flatten_ids,
                width, height, grad_mode=True):
        # `grad_mode`: torch.is_grad_enabled() at the call site (always off in here, and
        # needs_input_grad ignores it): a render under no_grad keeps no weight tiles
        _C.require_cuda(colors, geom)
        ctx.sink = None
        ctx.rows_marked = False
        if (_direct(colors) and colors.is_leaf and colors.requires_grad
                and colors.dtype == torch.float32 and colors.is_contiguous()):
            ctx.sink = colors
        colors = _f32c(colors)
        N, D = colors.shape
        dev = colors.device
        ev_param = param_ready_events.pop(colors.data_ptr(), None) if param_ready_events else None
        lazy = lazy_owners.get(colors.data_ptr()) if lazy_owners else None
        if D > 32 and D % 4 != 0:
            raise ValueError("wide blend needs D % 4 == 0 (rasterization() pads for you)")
        bg = _f32c(background) if background is not None else None
        render = torch.empty(height, width, D, dtype=torch.float32, device=dev)
        alphas = torch.empty(height, width, dtype=torch.float32, device=dev)
        # frozen geometry + trainable features (the shipped training loop): keep the blend-weight
        # tiles of this forward so the backward is a streaming GEMM instead of a second tile walk
        need_geo = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        last_ids = None
        if need_geo or want_last_ids or not _C.lib.gags_blend_last_ids_optional(D):
            last_ids = torch.empty(height, width, dtype=torch.int32, device=dev)
        cache = None
        if (weight_cache and grad_mode and ctx.needs_input_grad[3] and not need_geo
                and _C.lib.gags_blend_cache_supported(D)):
            n_tiles = offsets.numel() - 1
            slots = int(_C.lib.gags_blend_cache_slots(flatten_ids.numel(), n_tiles))
            lease = _CacheLease(dev, slots, ((width + TILE - 1) // TILE) * ((height + 7) // 8))
            cache = lease.bufs
            cur = torch.cuda.current_stream(dev)
            # the gradient buffer's row flags (persistent .grad of a row-sparse optimiser), if any
            rf = None
            if ctx.sink is not None and ctx.sink.grad is not None and row_flags:
                rf = row_flags.get(ctx.sink.grad.data_ptr())
            if lazy is not None and rf is not None and lazy.flags is rf and two_pass_forward:
                # two passes: nothing before the catch-up reads a feature, so exactly the rows this
                # view blends are brought up to date — and the weights pass does not have to wait
                # for the previous optimiser step / exchange either
                _C.check(_C.lib.gags_blend_fwd_weights(
                    _C.ptr(geom), width, height, _C.ptr(offsets), _C.ptr(flatten_ids),
                    _C.ptr(alphas), _C.ptr(last_ids), _C.ptr(cache[0]), _C.ptr(cache[1]),
                    _C.ptr(cache[2]), _C.ptr(cache[3]), _C.stream_ptr()), "gags_blend_fwd_weights")
                _mark("fwd_weights")
                if ev_param is not None:
                    cur.wait_event(ev_param)
                    ev_param = None
                evs = sink_ready_events.pop(ctx.sink.grad.data_ptr(), None)
                if evs is not None:                  # flags / gradient still in use by the last step
                    cur.wait_event(evs)
                _C.check(_C.lib.gags_blend_cache_mark_rows(
                    width, height, _C.ptr(offsets), _C.ptr(cache[1]), _C.ptr(cache[2]),
                    _C.ptr(cache[3]), _C.ptr(rf.flags), _C.stream_ptr()), "gags_blend_cache_mark_rows")
                rf.dirty = True
                ctx.rows_marked = True
                lazy.catch_up(rf.flags)
                _mark("rows_catch_up")
                _C.check(_C.lib.gags_blend_fwd_from_cache(
                    _C.ptr(colors), D, _C.ptr(bg), width, height, _C.ptr(offsets), _C.ptr(cache[0]),
                    _C.ptr(cache[1]), _C.ptr(cache[2]), _C.ptr(cache[3]), _C.ptr(alphas),
                    _C.ptr(render), _C.stream_ptr()), "gags_blend_fwd_from_cache")
                _C.count_launch(2)
            else:
                if ev_param is not None:
                    cur.wait_event(ev_param)
                    ev_param = None
                if lazy is not None:
                    lazy.flush()
                _C.check(_C.lib.gags_blend_fwd_cached(
                    _C.ptr(geom), _C.ptr(colors), D, _C.ptr(bg), width, height, _C.ptr(offsets),
                    _C.ptr(flatten_ids), _C.ptr(render), _C.ptr(alphas), _C.ptr(last_ids),
                    _C.ptr(cache[0]), _C.ptr(cache[1]), _C.ptr(cache[2]), _C.ptr(cache[3]),
                    _C.stream_ptr()), "gags_blend_fwd_cached")
        else:
            if ev_param is not None:
                torch.cuda.current_stream(dev).wait_event(ev_param)
            if lazy is not None:
                lazy.flush()
            _C.check(_C.lib.gags_blend_fwd(_C.ptr(geom), _C.ptr(colors), D, _C.ptr(bg), width,
                                           height, _C.ptr(offsets), _C.ptr(flatten_ids),
                                           _C.ptr(render), _C.ptr(alphas), _C.ptr(last_ids),
                                           _C.stream_ptr()), "gags_blend_fwd")
        _C.count_launch((D + 255) // 256 if D > 32 else 1)
        ctx.dims = (width, height, D, N)
        # The backward accumulates into a zeroed [N, D] buffer (2 GB at config 3).  Zeroing it is a
        # pure HBM stream, the forward above is instruction-bound with its shared memory full: the
        # fill runs beside it on the side stream instead of in front of the backward.
        ctx.prezero = None
        pending = getattr(_tls, "pending_prezero", None)
        if cache is not None and pending is not None and pending[0].shape == (N, D):
            ctx.prezero = pending
        _tls.pending_prezero = None
        ctx.lease = lease if cache is not None else None
        _tls.last_cached_ctx = ctx if cache is not None else None
        ctx.save_for_backward(colors, bg, geom, offsets, flatten_ids, alphas, last_ids)
        if last_ids is not None:
            ctx.mark_non_differentiable(last_ids)
        return render, alphas, last_ids

    @staticmethod
    def backward(ctx, v_render, v_alphas, _vl):
        colors, bg, geom, offsets, flatten_ids, alphas, last_ids = ctx.saved_tensors
        cache = ctx.lease.bufs if ctx.lease is not None else None
        width, height, D, N = ctx.dims
        dev = colors.device
        need_geo = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        need_col = ctx.needs_input_grad[3]
        if v_render is None:
            v_render = torch.zeros(height, width, D, device=dev)
        v_render = _f32c(v_render)
        st = _C.stream_ptr()
        _mark("bwd_start")
        v_colors, sink = _take_grad_buffer(ctx, need_col, need_geo, N, D, dev)
        _mark("bwd_zero")
        v_m = v_c = v_o = v_bg = None
        if not need_geo:
            if need_col and cache is not None:
                _C.check(_C.lib.gags_blend_bwd_features_cached(
                    D, width, height, _C.ptr(offsets), _C.ptr(cache[0]), _C.ptr(cache[1]),
                    _C.ptr(cache[2]), _C.ptr(cache[3]), _C.ptr(v_render), _C.ptr(v_colors), st),
                    "gags_blend_bwd_features_cached")
                _C.count_launch((D + 255) // 256)
            elif need_col:
                # frozen geometry: the feature-only fast path
                _C.check(_C.lib.gags_blend_bwd_features(_C.ptr(geom), D, width, height,
                                                        _C.ptr(offsets), _C.ptr(flatten_ids),
                                                        _C.ptr(v_render), _C.ptr(v_colors), st),
                         "gags_blend_bwd_features")
                _C.count_launch((D + 255) // 256 if D > 32 else 1)
        else:
            v_m = torch.zeros(N, 2, device=dev)
            v_c = torch.zeros(N, 3, device=dev)
            v_o = torch.zeros(N, device=dev)
            va = _f32c(v_alphas) if v_alphas is not None else None
            _C.check(_C.lib.gags_blend_bwd_full(_C.ptr(geom), _C.ptr(colors), D, _C.ptr(bg), width,
                                                height, _C.ptr(offsets), _C.ptr(flatten_ids),
                                                _C.ptr(alphas), _C.ptr(last_ids), _C.ptr(v_render),
                                                _C.ptr(va), _C.ptr(v_m), _C.ptr(v_c), _C.ptr(v_o),
                                                _C.ptr(v_colors), st), "gags_blend_bwd_full")
            _C.count_launch(2)
        if need_col and not ctx.rows_marked:
            _mark_rows(v_colors, cache if not need_geo else None, offsets, width, height)
        _mark("blend_bwd")
        if ctx.needs_input_grad[4] and bg is not None:
            v_bg = (v_render * (1.0 - alphas)[..., None]).sum(dim=(0, 1))
        if sink is not None:
            if sink.grad is None:
                sink.grad = v_colors                 # first view of the step: adopt the buffer
            v_colors = None                          # already accumulated; nothing for autograd to add
        return v_m, v_c, v_o, v_colors, v_bg, None, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# public operator
# ------------------------------------------------------------------------------------------------
def rasterize_view(means, quats, scales, opacities, colors, viewmat, fx, fy, cx, cy, width: int,
                   height: int, background=None, sh_degree: Optional[int] = None,
                   render_mode: str = "RGB", eps2d: float = 0.3, near_plane: float = 0.01,
                   far_plane: float = 1e10, radius_clip: float = 0.0,
                   scaling_modifier: float = 1.0, flags: int = 0):
    """One view.  `flags` selects the fused activations (raw log-scales / logit opacities).
    Returns render [H,W,D'], alphas [H,W], info."""
    if render_mode not in ("RGB", "D", "ED", "RGB+D", "RGB+ED"):
        raise ValueError(f"unknown render_mode {render_mode}")
    tile_w = (width + TILE - 1) // TILE
    tile_h = (height + TILE - 1) // TILE
    if 32 + tile_bits(tile_w * tile_h) > 64:
        raise ValueError("image too large for the 64-bit intersection key")
    geo_in = (means, quats, scales, opacities)
    _early_prezero(colors, geo_in, sh_degree)
    use_side = (lookahead and stage_events is None and means.is_cuda and viewmat.is_cuda
                and not any(t.requires_grad for t in geo_in))
    if use_side:
        dev = means.device
        main = torch.cuda.current_stream(dev)
        ss = _side_state(dev)
        side = ss["stream"]
        if not _all_seen(ss, geo_in + (viewmat,)):
            side.wait_stream(main)
        elif ss["ev_prev"] is not None:
            side.wait_event(ss["ev_prev"])
        with torch.cuda.stream(side):
            cam, keep = make_camera(viewmat, fx, fy, cx, cy, width, height, eps2d, near_plane,
                                    far_plane, radius_clip, scaling_modifier, flags)
            radii, means2d, depths, conics, opac, tiles, geom = _Project.apply(
                means, quats, scales, opacities, cam, keep, tile_w, tile_h)
            # Projection, the scan and the n_isects readback run early (beside the previous view's
            # loss / backward): the host is not held up behind the main stream's queue.  The emit /
            # sort kernels are then ordered behind everything the main stream has queued by now (up
            # to the optimiser pass): measured (tools/trace_step.py), a sort that is runnable while
            # Adam streams costs Adam 0.5 ms, and beside the persistent backward it cannot get an SM.
            binned = bin_and_sort(means2d, radii, depths, tiles, tile_w, tile_h,
                                  after_count=lambda: side.wait_stream(main))
            ev = torch.cuda.Event()
            ev.record(side)
        main.wait_event(ev)
        for t in (radii, means2d, depths, conics, opac, tiles, geom, keep) + binned["_bases"]:
            if t is not None:
                t.record_stream(main)
    else:
        cam, keep = make_camera(viewmat, fx, fy, cx, cy, width, height, eps2d, near_plane,
                                far_plane, radius_clip, scaling_modifier, flags)
        _mark("start")
        radii, means2d, depths, conics, opac, tiles, geom = _Project.apply(
            means, quats, scales, opacities, cam, keep, tile_w, tile_h)
        _mark("project")
        binned = bin_and_sort(means2d.detach(), radii, depths.detach(), tiles, tile_w, tile_h)
        _mark("bin_sort")
    if sh_degree is None:
        cols = colors
        if cols.dim() != 2 or cols.shape[0] != means.shape[0]:
            raise ValueError("colors must be [N, D] when sh_degree is None")
    else:
        vm = viewmat.detach().to(torch.float32)
        campos = torch.linalg.inv(vm)[:3, 3].contiguous().to(means.device)
        cols = _SHColors.apply(int(sh_degree), means, campos, colors, radii)
    bg = background
    if render_mode in ("RGB+D", "RGB+ED"):
        cols = torch.cat([cols, depths[:, None]], dim=-1)
        if bg is not None:
            bg = torch.cat([bg, torch.zeros(1, device=bg.device, dtype=bg.dtype)])
    elif render_mode in ("D", "ED"):
        cols = depths[:, None]
        bg = None
    D = cols.shape[1]
    pad = (-D) % 4 if D > 32 else 0
    if pad:
        cols = torch.nn.functional.pad(cols, (0, pad))
        if bg is not None:
            bg = torch.nn.functional.pad(bg, (0, pad))
    # the [1,N,2] handle is what render() hands out as "viewspace_points"; blending through its
    # [0] view makes .retain_grad() on it behave as in the reference (:75-78)
    means2d_c = means2d.unsqueeze(0)
    render, alphas, last_ids = _Blend.apply(means2d_c[0], conics, opac, cols, bg, geom,
                                            binned["offsets"], binned["flatten_ids"], width, height,
                                            torch.is_grad_enabled())
    _mark("blend_fwd")
    last_ctx = getattr(_tls, "last_cached_ctx", None)
    if last_ctx is not None:
        if not pad and render_mode == "RGB":
            render._gags_fused = _FusedHandle(last_ctx, cols, binned["offsets"], render.data_ptr())
        _tls.last_cached_ctx = None
    if use_side:
        ss["ev_prev"] = torch.cuda.Event()
        ss["ev_prev"].record(main)
    if pad:
        render = render[..., :D]
    if render_mode in ("ED", "RGB+ED"):
        render = torch.cat([render[..., :-1],
                            render[..., -1:] / alphas.clamp(min=1e-10)[..., None]], dim=-1)
    info = dict(radii=radii, means2d=means2d_c, depths=depths, conics=conics, opacities=opac,
                tile_width=tile_w, tile_height=tile_h, tiles_per_gauss=tiles,
                isect_ids=binned["isect_ids"], flatten_ids=binned["flatten_ids"],
                isect_offsets=binned["offsets"][:-1].reshape(tile_h, tile_w),
                n_isects=binned["n_isects"], last_ids=last_ids, width=width, height=height,
                tile_size=TILE, n_cameras=1, geom=geom)
    return render, alphas, info


def rasterization(means, quats, scales, opacities, colors, viewmats, Ks, width: int, height: int,
                  near_plane: float = 0.01, far_plane: float = 1e10, radius_clip: float = 0.0,
                  eps2d: float = 0.3, sh_degree: Optional[int] = None, packed: bool = False,
                  tile_size: int = 16, backgrounds=None, render_mode: str = "RGB", **unused):
    """Same call shape as the reference's use of gsplat.rasterization
    (/root/reference/gaussian_renderer/__init__.py:56-70): activated inputs, viewmats [C,4,4],
    Ks [C,3,3], backgrounds [C,D]; returns (colors [C,H,W,D'], alphas [C,H,W,1], info) with
    info["radii"] [C,N] and info["means2d"] [C,N,2] — the two keys GAGS reads (:74,:76,:83)."""
    if tile_size != TILE:
        raise ValueError("tile_size must be 16")
    if packed:
        raise ValueError("packed=True is not supported (the reference passes packed=False)")
    if viewmats.dim() != 3 or Ks.dim() != 3 or viewmats.shape[0] != Ks.shape[0]:
        raise ValueError("viewmats [C,4,4] and Ks [C,3,3] expected")
    C = viewmats.shape[0]
    Kh = Ks.detach().cpu().tolist()
    outs, alps, infos = [], [], []
    for c in range(C):
        bg = backgrounds[c] if backgrounds is not None else None
        r, a, info = rasterize_view(means, quats, scales, opacities, colors, viewmats[c],
                                    Kh[c][0][0], Kh[c][1][1], Kh[c][0][2], Kh[c][1][2], width,
                                    height, bg, sh_degree, render_mode, eps2d, near_plane,
                                    far_plane, radius_clip)
        outs.append(r)
        alps.append(a[..., None])
        infos.append(info)
    info = dict(infos[0])
    if C == 1:
        # keep means2d as a differentiable view so that .retain_grad() works as in the reference
        info["radii"] = infos[0]["radii"][None]
        info["means2d"] = infos[0]["means2d"]
        info["depths"] = infos[0]["depths"][None]
        info["conics"] = infos[0]["conics"][None]
    else:
        for k in ("radii", "depths", "conics"):
            info[k] = torch.stack([i[k] for i in infos])
        info["means2d"] = torch.cat([i["means2d"] for i in infos])
    info["n_cameras"] = C
    return torch.stack(outs), torch.stack(alps), info
