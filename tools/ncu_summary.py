#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python tools/ncu_summary.py full gpurun_out/prof_blend.ncu-rep profiles/r01_blend_full.md
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "smsp__warps_active.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.pct",
]


def short(name: str) -> str:
    name = name.replace("void ", "").replace("<unnamed>::", "")
    return name.split("(")[0]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        u = r[ui]
        ms = v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        k = short(r[ki])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}): gpu__time_duration.sum per kernel, "
                "--clock-control none, serialised cold-cache replays — compare SHARES\n\n")
        f.write("| kernel | launches | total ms | avg ms | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {ms:.3f} | {ms / n:.4f} | {100 * ms / total:.1f}% |\n")
        f.write(f"\ntotal {total:.3f} ms over {sum(a[0] for a in agg.values())} launches\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src}); per launch\n\n")
        for r in rows[2:]:
            f.write(f"## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for m in KEEP:
                if m in hdr:
                    i = hdr.index(m)
                    f.write(f"| {m} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    # DRAM bytes per launch of the two blend kernels -> profiles/traffic.json (read by bench.py)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    for r in rows[2:]:
        name = short(r[hdr.index("Kernel Name")])
        # blend_fwd_tc<NATOM, V3, MODE>: MODE 1 = weights pass, 2 = blend pass, 0 = single pass
        key = None
        if "blend_fwd_pers" in name:
            key = "blend_fwd"                                   # the persistent blend pass
        elif "blend_fwd" in name:
            mode = name.rstrip("> ").split(",")[-1].strip() if "<" in name else "0"
            key = {"1": "fwd_weights", "2": "blend_fwd"}.get(mode, "blend_fwd_single")
        elif "blend_bwd" in name:
            key = "blend_bwd"
        if key is None:
            continue
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(r[i].replace(",", "")) * scale.get(units[i], 1.0)
        traffic[key] = int(tot)
    if traffic:
        import json, os
        with open(os.path.join(os.path.dirname(dst), "traffic.json"), "w") as f:
            json.dump(traffic, f)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
