#!/usr/bin/env python3
"""Generates tests/golden/full_hashes.json ON THE GPU BOX: the synthetic benchmark fields are made by the
device generator (the bytes the benches and GPU tests use), brought to the host and extracted by the CPU oracle in lean
mode; counts and SHA-256 of both streams are recorded.  Full-size cases (fbm512, gyroid1024) take minutes of CPU time --
which is why they are committed hashes and not live oracle runs in the GPU suite; the windows are the first cell layers.
Test infrastructure (uses oracle/).   usage: gen_golden_full.py [out.json] [case ...]"""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import sha  # noqa: E402
from oracle import oracle as O  # noqa: E402

# name: (field kind, size, seed, cell layers from z = 0 (None = all))
CASES = {
    "fbm512": (1, 512, 0x1505F00D, None),
    "fbm644_z32": (1, 644, 0x1505F00D, 32),
    "fbm812_z32": (1, 812, 0x1505F00D, 32),
    "fbm1024_z24": (1, 1024, 0x1505F00D, 24),
    "gyroid1024": (2, 1024, 0, None),
    "spheres2048_z64": (3, 2048, 0x5EEDBA11, 64),
}


def device_field(kind, size, seed, n_layers):
    import torch
    from isosurface_b200 import _lib
    t = torch.empty(n_layers * size * size, dtype=torch.float32, device="cuda:0")
    _lib.check(_lib.load().isomc_synth_field(0, kind, size, seed, 0, n_layers, C.c_void_p(t.data_ptr())))
    return t.cpu().numpy().reshape(n_layers, size, size)


def main():
    out_path = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "tests" / "golden" / "full_hashes.json"
    names = sys.argv[2:] or list(CASES)
    out = json.loads(out_path.read_text()) if out_path.exists() else {}
    for name in names:
        kind, size, seed, zc = CASES[name]
        z = size if zc is None else zc
        f = device_field(kind, size, seed, z + 1)
        t0 = time.time()
        xyz, idx, act = O.extract_grid(size, f, z, O.LEAN)
        out[name] = {"kind": kind, "size": size, "seed": seed, "z_cells": z, "active_cells": int(act), "vertices": len(xyz) // 3,
                     "triangles": len(idx) // 3, "sha_v": sha(xyz, "<f4"), "sha_i": sha(idx, "<u4"),
                     "field_sha": sha(f, "<f4"), "oracle_seconds": round(time.time() - t0, 1)}
        print(name, out[name], flush=True)
        out_path.write_text(json.dumps(out, indent=1) + "\n")


if __name__ == "__main__":
    main()
