#!/usr/bin/env python3
"""Randomised parity stress on the GPU: random sizes, densities, fields and slab counts against the CPU oracle.
   python tools/gpu_stress.py [n_cases] [seed]        (test infrastructure; uses oracle/)"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import torch  # noqa: E402

import isosurface_b200 as iso  # noqa: E402
from helpers import mesh_diff  # noqa: E402
from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers  # noqa: E402
from oracle import oracle as O  # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
bad = 0
t0 = time.time()
for case in range(n_cases):
    n = int(rng.choice([2, 3, 4, 5, 31, 32, 33, 34, 63, 64, 65, 66, 95, 96, 97, 100, 127, 128, 129, 130, 160]))
    if rng.random() < 0.4:
        n = int(rng.integers(2, 140))
    kind = rng.integers(0, 4)
    shape = (n + 1, n, n)
    if kind == 0:      # white noise with a random threshold: density from very sparse to every cell active
        g = rng.standard_normal(shape).astype(np.float32) + np.float32(rng.uniform(-2.5, 2.5))
    elif kind == 1:    # smooth random field
        z, y, x = np.meshgrid(np.arange(n + 1), np.arange(n), np.arange(n), indexing="ij")
        k = rng.uniform(0.05, 0.9, size=3)
        g = (np.sin(k[0] * x + rng.uniform(0, 6)) * np.cos(k[1] * y) + np.sin(k[2] * z + rng.uniform(0, 6))).astype(np.float32)
    elif kind == 2:    # blocky field with many exact zeros (inside) and ties
        g = rng.integers(-1, 2, size=shape).astype(np.float32)
    else:              # planes aligned with the lattice boundaries (stress the low/high faces)
        g = np.ones(shape, np.float32)
        g[: rng.integers(0, 3)] = -1
        g[:, : rng.integers(0, 3)] = -1
        g[:, :, : rng.integers(0, 3)] = -1
        g[-int(rng.integers(0, 3)) or None:] *= -1
        g += (rng.standard_normal(shape) * 0.1).astype(np.float32)
    oxyz, oidx, oact = O.extract_grid(n, g)
    mc = iso.MarchingCubes(n)
    nv, nt, na = mc.extract_device(iso.DenseGrid(g))
    xyz, idx = mc.copy_out()
    mc.close()
    msg = mesh_diff(xyz, idx, oxyz, oidx)
    if msg == "" and na != oact:
        msg = "active cells %d != %d" % (na, oact)
    world = int(rng.integers(2, 6))
    if msg == "" and n >= world * 2 and rng.random() < 0.5:
        t = torch.from_numpy(g).cuda()
        slabs = [SlabMarchingCubes(n, r, world) for r in range(world)]
        ptrs = [t.data_ptr() + 4 * slab_sample_layers(n, r, world)[0] * n * n for r in range(world)]
        totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
        parts = []
        for r, s in enumerate(slabs):
            s.extract(ptrs[r], gathered=totals)
            parts.append(s.copy_out())
            s.close()
        msg = mesh_diff(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), oxyz, oidx)
        if msg:
            msg = "slabs(world=%d): " % world + msg
    if msg:
        bad += 1
        print("CASE %d n=%d kind=%d FAILED: %s" % (case, n, kind, msg), flush=True)
print("%d cases, %d failures, %.1f s" % (n_cases, bad, time.time() - t0))
sys.exit(1 if bad else 0)
