#!/usr/bin/env python3
"""Small extracts for compute-sanitizer (memcheck / racecheck / initcheck): implicit source, dense grid incl. a
white-noise field (dense multi-pass path of k_emit), a slab with a ghost layer."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import isosurface_b200 as iso  # noqa: E402
from helpers import iso_source  # noqa: E402
from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers  # noqa: E402

mc = iso.MarchingCubes(40)
print("csgA", mc.extract_device(iso.Sampler(iso_source("csgA"))))
rng = np.random.default_rng(3)
g = rng.standard_normal((41, 40, 40)).astype(np.float32)
print("noise", mc.extract_device(iso.DenseGrid(g)))
mc.close()
mc = iso.MarchingCubes(70)
g = rng.standard_normal((71, 70, 70)).astype(np.float32)
print("noise70", mc.extract_device(iso.DenseGrid(g)))
mc.close()
import torch  # noqa: E402
size, world = 36, 3
t = torch.from_numpy(rng.standard_normal((size + 1, size, size)).astype(np.float32)).cuda()
slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
for r, s in enumerate(slabs):
    s.extract(ptrs[r], gathered=totals)
    print("slab", r, [len(a) for a in s.copy_out()])
    s.close()

# round 2: repeated extracts (CUDA-graph replay of the launch sequence), batched chunks, the peer-memory totals exchange
import ctypes as C  # noqa: E402
from isosurface_b200 import _lib  # noqa: E402
mc = iso.MarchingCubes(40)
for _ in range(3):
    print("graph replay", mc.extract_device(iso.Sampler(iso_source("torus"))))
mc.close()
drv = iso.BatchedMarchingCubes(24, n_chunks=6)
srcs = [iso.Sampler(iso.Translate((0.3 + 0.08 * i, 0.5, 0.5), iso.Sphere(0.2))) for i in range(9)]
print("batch", [len(m[1]) // 3 for m in drv.extract_many(srcs)])
drv.close()
# dense chunks through the same stacked-lattice handle: a device batch (used in place) and a partly filled host batch
gdrv = iso.BatchedMarchingCubes(20, n_chunks=4)
lat = np.random.default_rng(5).standard_normal((4, 21, 20, 20)).astype(np.float32)
print("batch grids", [int(v) for v in gdrv.extract_grids(torch.from_numpy(lat).cuda())[2]],
      [int(v) for v in gdrv.extract_grids(lat[:3])[2]])
gdrv.close()
ddrv = iso.BatchedMarchingCubes(24, n_chunks=4, distance="directed")  # MarchingCubes<Directed> per chunk, one padding lattice
print("batch directed", [int(v) for v in ddrv.extract_batch(srcs[:3])[2]])
ddrv.close()
lib = _lib.load()
slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
boxes = (C.c_void_p * world)()
for r, s in enumerate(slabs):
    b = C.c_void_p()
    _lib.check(lib.isomc_slab_mailbox(s._h, C.byref(b)), s._h)
    boxes[r] = b.value
for r, s in enumerate(slabs):
    _lib.check(lib.isomc_slab_connect(s._h, r, world, boxes), s._h)
for step in range(3):
    for r, s in enumerate(slabs):
        _lib.check(lib.isomc_slab_count_grid_device(s._h, C.c_void_p(ptrs[r])), s._h)
        _lib.check(lib.isomc_slab_enqueue_emit_exchanged(s._h), s._h)
    for s in slabs:
        _lib.check(lib.isomc_finish(s._h), s._h)
print("exchanged", [[len(a) for a in s.copy_out()] for s in slabs])
for s in slabs:
    s.close()
