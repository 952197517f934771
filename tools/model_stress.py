#!/usr/bin/env python3
"""Randomised stress of the HOST model of the list kernels (tests/list_model.cu) against the oracle: random sizes,
densities, warp counts and execution orders, thin wide lattices, slab pairs.  python tools/model_stress.py [cases] [seed]"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import mesh_diff  # noqa: E402
from oracle import oracle as O  # noqa: E402
import test_list_model as T  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
lib = T.model.__wrapped__() if hasattr(T.model, "__wrapped__") else None
if lib is None:
    import ctypes as C
    import subprocess
    if not T.SO.exists():
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-arch=sm_100a", "-DISOMC_HOST_MODEL", "-Xcompiler", "-ffp-contract=off,-fPIC,-fno-fast-math",
                        "-shared", "-o", str(T.SO), str(T.SRC)], check=True)
    lib = C.CDLL(str(T.SO))
    lib.list_model_extract.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
fails, t0 = 0, time.time()
for case in range(n_cases):
    kind = rng.integers(0, 4)
    if kind == 3:  # thin window of a wide lattice (rows of > 32 / > 64 segments)
        size, zc = int(rng.choice([1030, 1060, 1100, 2080, 2200])), 1
    else:
        size, zc = int(rng.choice([2, 3, 5, 17, 31, 32, 33, 34, 48, 63, 64, 65, 66, 70, 97])), None
    zl = (zc if zc else size) + 1
    f = rng.standard_normal((zl, size, size)).astype(np.float32)
    if kind == 1:  # threshold: density anywhere between empty and full
        f += np.float32(rng.uniform(-3, 3))
    if kind == 2:  # smooth, with exact zeros
        f = np.round(np.cumsum(np.cumsum(f, axis=2), axis=1) / 4).astype(np.float32)
    if kind == 3:
        lo, hi = sorted(rng.integers(0, size, 2))
        f[:, :, lo:hi] = np.abs(f[:, :, lo:hi]) * (1 if rng.random() < 0.5 else -1)
        f[:, ::int(rng.integers(2, 9)), :] = 1.0
    nw = int(rng.integers(1, 40))
    oxyz, oidx, oact = O.extract_grid(size, f, z_cells=zc)
    rc, xyz, idx, tot = T.run_model(lib, size, f, z_end=zc, n_warps=nw, seed=case, cap_blocks=70000)
    msg = mesh_diff(xyz, idx, oxyz, oidx) if rc == 0 else "rc=%d" % rc
    if msg or tot[3] != oact:
        fails += 1
        print("case %d kind %d size %d warps %d: %s (active %d vs %d)" % (case, kind, size, nw, msg, tot[3], oact), flush=True)
print("%d cases, %d failures, %.1f s" % (n_cases, fails, time.time() - t0))
sys.exit(1 if fails else 0)
