"""In-tree build of the C-ABI library (nvcc, sm_100a only).  `python -m gags_b200.build`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libgags_b200.so")
SOURCES = ["api.cu", "project.cu", "tiles.cu", "tile_buckets.cu", "sh.cu", "blend_fwd.cu", "blend_fwd_tc.cu", "blend_bwd.cu", "blend_bwd_tc.cu", "blend_bwd_cached.cu",
           "blend_bwd_geom.cu", "train_ops.cu", "pixel_losses.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; gags_b200 has no non-CUDA path")
    return exe


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(HERE, "..", "include", "gags_b200.h"))
    return max(os.path.getmtime(f) for f in files)


def build(force: bool = False, verbose: bool = False, timing: bool = False) -> str:
    """timing=True builds csrc/libgags_b200_dbg.so with -DGAGS_TC_TIMING (tools/tc_timeline.py)."""
    lib = LIB.replace(".so", "_dbg.so") if timing else LIB
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= _deps_mtime():
        return lib
    nvcc = _nvcc()
    objdir = os.path.join(CSRC, "build_dbg" if timing else "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src: str) -> str:
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if timing:
            cmd.insert(1, "-DGAGS_TC_TIMING")
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", lib, *objs, "-gencode",
                        "arch=compute_100a,code=sm_100a"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                timing="--timing" in sys.argv))
