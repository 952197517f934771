"""Builds libisomc_b200.so (in-tree) with nvcc for sm_100a.  Works without a GPU (cross-compile)."""
import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libisomc_b200.so"
SOURCES = ["isomc_kernels.cu", "isomc_list_kernels.cu", "isomc_tile_kernels.cu", "isomc_points.cu", "isomc_api.cu", "isomc_sharded.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false",  # the reference (rustc) never contracts a*b+c; keep sample signs / vertex bits identical
    "-Xcompiler", "-fPIC", "-shared", "-ldl",
]


def nvcc_path():
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("nvcc not found (set NVCC=...)")
    return cand


def is_stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [PKG.parent / "include" / "isomc.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force=False, verbose=False):
    if not force and not is_stale():
        return LIB
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", str(LIB)] + [str(CSRC / s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    build_library(force=True, verbose="-v" in sys.argv)
    print(LIB)
