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
