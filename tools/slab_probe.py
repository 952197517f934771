#!/usr/bin/env python3
"""Per-kernel times of ONE rank's slab of a sharded extract on a single GPU (what a rank of an N-GPU run does):
   python tools/slab_probe.py WORKLOAD WORLD RANK"""
import ctypes as C
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402
import bench  # noqa: E402
from isosurface_b200 import _lib  # noqa: E402

wl, world, rank = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
lib = _lib.load()
env = {"lib": lib, "world": world, "rank": rank, "local": 0, "stream": torch.cuda.Stream(device=0)}
r = bench.Runner.__new__(bench.Runner)
# a Runner without the collective: count, then emit with bases taken from this rank alone (timing only)
r.env, r.wl, r.torch, r._lib, r.lib = env, wl, torch, _lib, lib
r.world, r.rank, r.local = 1, 0, 0
size, kind, field, seed = bench.WORKLOADS[wl]
z0, z1 = bench.slab_range(size, rank, world)
ghost = 1 if z0 > 0 else 0
h = C.c_void_p()
_lib.check(lib.isomc_slab_create(size, z0, z1, 0, C.byref(h)))
grid = bench.make_field(lib, torch, 0, wl, z0 - ghost, (z1 - z0) + ghost + 1)
_lib.check(lib.isomc_set_profiling(h, 1), h)
out = []
for it in range(6):
    _lib.check(lib.isomc_slab_count_grid_device(h, C.c_void_p(grid.data_ptr())), h)
    _lib.check(lib.isomc_slab_emit(h, 0, 0), h)
    st = _lib.Stats()
    _lib.check(lib.isomc_stats_get(h, C.byref(st)), h)
    out.append((st.ms_sign, st.ms_count, st.ms_scan, st.ms_emit, st.ms_total))
import numpy as np  # noqa: E402
m = np.array(out[2:]).mean(axis=0)
print(json.dumps({"workload": wl, "world": world, "rank": rank, "k_sign": m[0], "k_count_list": m[1], "k_scan_rows": m[2], "emit": m[3],
                  "total": m[4], "vertices": int(st.n_vertices)}))
