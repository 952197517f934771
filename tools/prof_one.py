#!/usr/bin/env python3
"""Runs a few extracts of one bench workload (for ncu):  python tools/prof_one.py fbm512 [n_extracts]"""
import ctypes as C
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

import bench  # noqa: E402
from isosurface_b200 import _lib  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "fbm512"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib = _lib.load()
size, kind, field, seed = bench.WORKLOADS[wl]
h = C.c_void_p()
_lib.check(lib.isomc_create(size, 0, C.byref(h)))
if kind == "grid":
    grid = bench.make_field(lib, torch, 0, wl, 0, size + 1)
    for _ in range(n):
        _lib.check(lib.isomc_extract_grid_device(h, C.c_void_p(grid.data_ptr())), h)
else:
    sys.path.insert(0, str(ROOT / "tests"))
    from helpers import iso_source
    from isosurface_b200.source import encode_program
    prog = encode_program(iso_source(field))
    for _ in range(n):
        _lib.check(lib.isomc_extract_sdf(h, prog.ctypes.data, len(prog)), h)
st = _lib.Stats()
_lib.check(lib.isomc_stats_get(h, C.byref(st)), h)
print(wl, "V", st.n_vertices, "T", st.n_triangles, "launches/extract", st.kernel_launches)
