#!/usr/bin/env python3
"""Chunk throughput of small extracts (C1a: 32^3 sphere, C1b: 128^3 torus): one handle at a time vs. the multi-chunk
driver with n handles in flight:  python tools/bench_chunks.py  -> one JSON line per case"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import isosurface_b200 as iso  # noqa: E402
from helpers import iso_source  # noqa: E402

for name, size, shape in (("sphere32", 32, "sphere03"), ("torus128", 128, "torus_origin")):
    src = iso.Sampler(iso_source(shape))
    n = 256
    for k in (1, 2, 4, 8, 16):
        drv = iso.ChunkedMarchingCubes(size, n_inflight=k)
        drv.extract_many([src] * 32)  # warm-up: buffers sized, list sized
        t0 = time.perf_counter()
        drv.extract_many([src] * n, deliver=lambda i, xyz, idx: None)
        dt = time.perf_counter() - t0
        drv.close()
        print(json.dumps({"workload": name, "n_inflight": k, "chunks": n, "us_per_chunk": 1e6 * dt / n, "chunks_per_s": n / dt,
                          "note": "wall clock incl. enqueue, finish and copy-out of every chunk's mesh to host"}), flush=True)

# batched chunks: B programs through ONE kernel sequence (isomc_extract_sdf_batch), one size read-back, one copy-out
for name, size, shape in (("sphere32", 32, "sphere03"), ("torus128", 128, "torus_origin")):
    src = iso.Sampler(iso_source(shape))
    n = 512
    for B in (8, 32, 128):
        if B * (size + 1) * (size - 1) >= 1 << 26:
            continue
        drv = iso.BatchedMarchingCubes(size, n_chunks=B)
        drv.extract_many([src] * (2 * B))  # warm-up: buffers sized, list sized
        t0 = time.perf_counter()
        drv.extract_many([src] * n, deliver=lambda i, xyz, idx: None)
        dt = time.perf_counter() - t0
        # extraction alone (no copy-out): what the kernel sequence + one synchronisation cost per chunk
        import numpy as np
        from isosurface_b200 import _lib
        from isosurface_b200.source import encode_program
        progs = [encode_program(src.source)] * B
        flat = np.concatenate(progs)
        nn = np.asarray([len(p) for p in progs], np.uint32)
        t1 = time.perf_counter()
        for _ in range(n // B):
            _lib.check(drv._lib.isomc_extract_sdf_batch(drv._h, flat.ctypes.data, nn.ctypes.data, B), drv._h)
        dx = time.perf_counter() - t1
        drv.close()
        print(json.dumps({"workload": name, "batch": B, "chunks": n, "us_per_chunk": 1e6 * dt / n, "chunks_per_s": n / dt,
                          "us_per_chunk_extract_only": 1e6 * dx / (n // B * B),
                          "note": "isomc_extract_sdf_batch; wall clock incl. one copy-out per batch and the per-chunk split on the host; "
                                  "extract_only = kernel sequence + size read-back, mesh left on the device"}), flush=True)
