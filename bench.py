#!/usr/bin/env python
"""bench.py — views/sec of the GAGS distillation step on synthetic Gaussians (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # product arm (sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K ...  # reference arm: the CPU oracle port
  torchrun ... bench.py --gpus N ...                       # N > 1: one rank per GPU, view-parallel

One step = one optimiser step of the train.py loop (/root/reference/train.py:109-228) with frozen
geometry on the config-3 shape (N=2M Gaussians, 1080x1920, D=256): for each of `views_per_step`
views per rank (default 1 at every N)  render() -> L1 against the view's emb[seg] target fused with
the backward to the per-Gaussian features (l1_backward_fused);  then Adam on the feature table —
at N>1 fused with the gradient all-reduce and the parameter all-gather into one NVLink peer-memory
kernel (parallel.PeerAdam; --nccl-allreduce = NCCL all-reduce + local Adam, also the fallback).
value = views/sec over the whole job (all ranks), inputs resident in HBM.
e2e   = the same loop with the per-view training target (segment map + embedding table, the
        inputs of the reference's read_sam_clip_feature) copied from pinned host memory each view
        and the loss read back to the host each step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "views/sec fwd+bwd (N=2M Gaussians, 1080p, D=256)"      # config 3, the headline


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gags_b200", choices=["gags_b200", "reference"])
    ap.add_argument("--config", type=int, default=3, help="BASELINE.json config id (3 = headline)")
    ap.add_argument("--views-per-step", type=int, default=None,
                    help="views rendered per rank between optimiser steps (default 1; 4 when N>1)")
    ap.add_argument("--n", type=int, default=None, help="override N (debug only; invalidates value)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--nccl-allreduce", action="store_true",
                    help="N>1: NCCL all-reduce + local Adam instead of the fused peer-memory step")
    ap.add_argument("--unfused-loss", action="store_true",
                    help="separate L1 kernel + autograd backward instead of l1_backward_fused")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML during the timed region."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag = index, [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
                 nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons)}


# ------------------------------------------------------------------------------------------------
def reference_arm(args):
    """The reference's implementation of this path (gsplat) is CUDA-only and absent; its CPU
    stand-in is the oracle port, timed on all host cores on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    from gags_b200.synthetic import CONFIGS, config_scene
    from oracle.cpu_baseline import time_view
    n, h, w, d = CONFIGS[args.config]
    scene = config_scene(args.config)
    times, info = [], None
    for i in range(args.warmup + args.steps):
        info = time_view(scene, scene.cameras[i % len(scene.cameras)], d, n_tiles=24, seed=i)
        if i >= args.warmup:
            times.append(info["seconds_per_view"])
        if sum(times) > 150:
            break
    sec = sum(times) / len(times)
    line = {"impl": "reference", "metric": METRIC, "value": 1.0 / sec, "unit": "views/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"config{args.config}: N={n}, {h}x{w}, D={d}, fwd+bwd_feat, "
                                   "CPU oracle port (gsplat itself is CUDA-only and not installed)"},
            "cpu_baseline": {"value": 1.0 / sec, "unit": "views/s", "cores": info["cores"],
                             "kind": "port", "sample": info["sample"]},
            "e2e": {"value": 1.0 / sec, "unit": "views/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
def main():
    args = parse()
    if args.impl == "reference":
        return reference_arm(args)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=gags_b200) needs a CUDA device: there is no CPU fallback")
    import torch.distributed as dist
    from gags_b200 import _C, parallel, rasterization as R
    from gags_b200.arguments import OptimizationParams
    from gags_b200.gaussian_renderer import render
    from gags_b200.scene import GaussianModel
    from gags_b200.synthetic import CONFIGS, config_scene
    from gags_b200.utils.loss_utils import l1_backward_fused, l1_loss_segmap_fused

    rank, world, local = parallel.init_from_env("nccl")
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # one view per rank per optimiser step at every N (the train.py loop on each rank): per-GPU work
    # is the same at 1, 2, 4 and 8 GPUs; --views-per-step k accumulates k views before the exchange
    kviews = args.views_per_step or 1

    n, H, W, D = CONFIGS[args.config]
    scene = config_scene(args.config)
    if args.n:
        n = args.n
        for f in ("xyz", "scaling", "rotation", "opacity", "features_dc", "features_rest",
                  "semantic_feature"):
            setattr(scene, f, getattr(scene, f)[:n].contiguous())
    pc = GaussianModel(3, device=dev)
    pc.create_from_tensors(scene.xyz, scene.scaling, scene.rotation, scene.opacity,
                           scene.features_dc, scene.features_rest, scene.semantic_feature)
    pc.training_setup(OptimizationParams(), fused_optimizer=True)
    cams = [c.to(dev) for c in scene.cameras]
    n_views = len(cams)
    # the cameras are loaded once and never written again (scene/cameras.py:58): tell the renderer,
    # so its side-stream geometry stage need not wait for the main stream the first time it meets one
    R.register_static(*[c.world_view_transform for c in cams])
    bg = torch.zeros(3, device=dev)

    # Targets: a compact (segment map, embedding table) pair per target, like the reference's
    # (seg_map, img_embed) inputs to read_sam_clip_feature (scene/dataset_readers.py:54-121).
    n_targets, n_seg = 2, 256
    g = torch.Generator().manual_seed(4321)
    seg_host = [torch.randint(0, n_seg, (H // 8 + 1, W // 8 + 1), generator=g, dtype=torch.int32)
                .repeat_interleave(8, 0).repeat_interleave(8, 1)[:H, :W].contiguous().pin_memory()
                for _ in range(n_targets)]
    emb_host = [(0.1 * torch.randn(n_seg, D, generator=g)).pin_memory() for _ in range(n_targets)]

    # the targets stay in that compact form on the device as well: the fused loss gathers
    # emb[seg] on the fly (the reference materialises the dense map every iteration, train.py:162)
    targets_dev = [(s.to(dev), e.to(dev)) for s, e in zip(seg_host, emb_host)]

    # k views per optimiser step: let the backward reduce straight into .grad (see rasterization.py)
    R.direct_grad_accumulation = kviews > 1

    # Multi-GPU gradient exchange: all-reduce + Adam + parameter all-gather fused into one kernel
    # over NVLink peer memory (parallel.PeerAdam); --nccl-allreduce selects the baseline (NCCL
    # all-reduce pipelined with the full local Adam pass), which is also the fallback when
    # symmetric memory cannot be set up on this box.
    peer, exchange = None, "none"
    if world > 1:
        exchange = "nccl-allreduce + local Adam"
        if not args.nccl_allreduce:
            try:
                grp = next(g for g in pc.optimizer.param_groups
                           if any(q is pc._semantic_feature for q in g["params"]))
                peer = parallel.PeerAdam(pc._semantic_feature, lr=grp["lr"], betas=grp["betas"],
                                         eps=grp["eps"])
                exchange = ("peer-memory fused all-reduce + sharded Adam + all-gather (one kernel, "
                            + ("NVLS multimem" if peer.multicast else "unicast P2P") + ")")
            except Exception as e:                       # noqa: BLE001 - reported, NCCL path used
                print(f"[bench] PeerAdam unavailable ({type(e).__name__}: {e}); using NCCL",
                      file=sys.stderr)
                peer = None
    if peer is not None:
        R.direct_grad_accumulation = True

    def one_view(cam, target, target_ready=None):
        pkg = render(cam, pc, None, bg)                      # feature_mode=True (default)
        if target_ready is not None:                         # the target's H2D copy (copy stream)
            torch.cuda.current_stream(dev).wait_event(target_ready)
        if args.unfused_loss:
            loss = l1_loss_segmap_fused(pkg["render"], target[0], target[1])
            loss.backward()
            return loss
        # loss + loss.backward() as one call: the L1 gradient is formed inside the feature backward
        return l1_backward_fused(pkg["render"], target[0], target[1])

    def opt_step():
        if peer is not None:
            peer.step()                                      # re-zeroes its persistent .grad itself
            return
        if world > 1:
            parallel.allreduce_and_step(pc.optimizer, pc._semantic_feature, world)
        else:
            pc.optimizer.step()
        pc.optimizer.zero_grad(set_to_none=True)

    def step_resident(step):
        loss = None
        for v in parallel.views_for_rank(step, rank, world, kviews, n_views):
            loss = one_view(cams[v], targets_dev[v % n_targets])
        opt_step()
        return loss

    copy_stream = torch.cuda.Stream(device=dev)

    def step_e2e(step):
        loss = None
        h2d = 0
        for v in parallel.views_for_rank(step, rank, world, kviews, n_views):
            cam = cams[v]                 # cameras live on the device (scene/cameras.py:58)
            # this view's target travels on a copy stream while the view renders; the loss waits
            with torch.cuda.stream(copy_stream):
                seg = seg_host[v % n_targets].to(dev, non_blocking=True)
                emb = emb_host[v % n_targets].to(dev, non_blocking=True)
                ready = torch.cuda.Event()
                ready.record(copy_stream)
            main = torch.cuda.current_stream(dev)
            seg.record_stream(main)
            emb.record_stream(main)
            h2d += seg.numel() * 4 + emb.numel() * 4
            loss = one_view(cam, (seg, emb), ready)
        opt_step()
        return float(loss.item()), h2d                       # D2H read of the step's loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for i in range(warmup):
            fn(i)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local)
        sampler.start()
        l0 = _C.launches()
        a0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
        marks = []
        e0.record()
        last = None
        for i in range(steps):
            last = fn(warmup + i)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        if peer is not None:
            peer.synchronize()        # the last step's exchange runs on its own stream: drain it
        e1.record()
        barrier()
        sampler.stop_flag = True
        sampler.join()
        ms = parallel.max_over_ranks(e0.elapsed_time(e1), world, dev)
        # diagnostics only: the slowest single step and the cudaMallocs that landed in the region
        per = [a.elapsed_time(b) for a, b in zip([e0] + marks[:-1], marks)]
        timed.diag = {"max_step_ms": round(max(per), 3), "min_step_ms": round(min(per), 3),
                      "device_allocs": torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - a0}
        return ms, last, sampler.summary(), _C.launches() - l0

    # Untimed priming (like the reference's scene loading): the grow-only workspaces (intersection
    # buffers, weight-tile cache) reach their steady size, so that no cudaMalloc — a device-wide
    # sync — lands in the timed region.  The W warm-up and K timed steps follow as the contract says.
    priming = 6
    for i in range(priming):
        step_resident(10_000 + i)
    ms, _, clocks, launches = timed(step_resident, args.steps, args.warmup)
    alloc_stats = {"reserved_GB": round(torch.cuda.max_memory_reserved(dev) / 1e9, 2)}
    alloc_stats.update(timed.diag)
    views_total = args.steps * kviews * world
    value = views_total / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        ms_e, last, _, _ = timed(step_e2e, max(3, args.steps // 2), 2)
        steps_e = max(3, args.steps // 2)
        e2e = {"value": steps_e * kviews * world / (ms_e * 1e-3), "unit": "views/s",
               "h2d_bytes_per_step": int(last[1]), "d2h_bytes_per_step": 4}

    if peer is not None:
        peer.synchronize()
        torch.cuda.synchronize(dev)
    # ---- per-stage device times for the roofline (rank 0, single views, CUDA events) -------------
    stage_ms, stats, roofline = {}, {}, None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_kind = "measured" if peaks else "fallback"
        acc = {}
        reps = 5
        for i in range(-2, reps):                 # two unrecorded passes: this single-stream,
            R.stage_events = []                   # instrumented path has its own allocation pattern
            cam = cams[(7 * i) % n_views]
            pkg = render(cam, pc, None, bg)
            R._mark("loss_start")
            if args.unfused_loss:
                loss = l1_loss_segmap_fused(pkg["render"], targets_dev[0][0], targets_dev[0][1])
                R._mark("loss")
                loss.backward()
            else:
                loss = l1_backward_fused(pkg["render"], targets_dev[0][0], targets_dev[0][1])
            R._mark("backward_end")
            pc.optimizer.step()
            R._mark("adam")
            pc.optimizer.zero_grad(set_to_none=True)
            torch.cuda.synchronize()
            ev = R.stage_events
            R.stage_events = None
            if i < 0:
                continue
            for (n0, a), (n1, b) in zip(ev[:-1], ev[1:]):
                acc.setdefault(n1, []).append(a.elapsed_time(b))
            if i == 0:
                radii = pkg["radii"]
                stats["n_visible"] = int((radii > 0).sum())
        stage_ms = {k: sum(v) / len(v) for k, v in acc.items()}
        nv = stats.get("n_visible", n)
        fwd_bytes = nv * 4 * D + H * W * (4 * D + 8)
        bwd_bytes = H * W * (4 * D + 8) + nv * 4 * D
        kname = "blend_fwd" if stage_ms.get("blend_fwd", 0) >= stage_ms.get("blend_bwd", 0) \
            else "blend_bwd"
        kbytes = fwd_bytes if kname == "blend_fwd" else bwd_bytes
        ach = kbytes / (stage_ms[kname] * 1e-3) / 1e9
        # DRAM bytes per launch of that kernel from the committed `ncu --set full` capture
        # (profiles/traffic.json, written by tools/ncu_summary.py); config 3 only
        traffic = None
        try:
            if args.config in (3, 4):
                traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[kname]
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": kname, "achieved": ach, "peak": hbm_peak,
                    "unit": "GB/s", "frac": ach / hbm_peak, "peak_kind": peak_kind,
                    "traffic": traffic, "algorithmic_bytes": kbytes,
                    "avg_launch_ms": stage_ms[kname]}
        step_bytes = n * (44 + 4 * D) + H * W * (4 * D + 4) + H * W * (4 * D + 8) + n * (4 * D + 24)
        stats["step_hbm_frac_of_8TBps"] = step_bytes / (ms * 1e-3 / (args.steps * kviews)) / 8e12
        stats.update(alloc_stats)
        if peer is not None and peer.timing_summary() is not None:
            stats["peer_step_ms"] = peer.timing_summary()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_baseline import time_view
        t0 = time.time()
        info = time_view(scene, scene.cameras[0], D, n_tiles=32)
        cpu = {"value": info["views_per_s"], "unit": "views/s", "cores": info["cores"],
               "kind": "port", "sample": info["sample"],
               "wall_s": round(time.time() - t0, 1)}
        stats["n_isects"] = info["n_isects"]

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "views/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"config{args.config}: N={n}, {H}x{W}, D={D}, render + fused "
                                       "L1 vs emb[seg] target + feature backward + fused Adam "
                                       "(frozen geometry)",
                           "priming_steps": priming,
                           "views_per_step_per_gpu": kviews, "parallelism": f"view-dp{world}",
                           "grad_exchange": exchange,
                           "l2": "inputs (2 GB feature table, 2 GB raster) exceed the 126 MB L2",
                           "optimizer_in_timed_region": True},
                "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu, "stage_ms": stage_ms, "stats": stats}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
