#!/usr/bin/env python3
"""First-contact GPU diagnostic: runs a ladder of cases through the C ABI and compares every stage
with the CPU oracle, printing details on the first mismatch.  Test infrastructure (uses oracle/)."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import isosurface_b200 as iso  # noqa: E402
from helpers import iso_source, mesh_diff, oracle_prog  # noqa: E402
from oracle import oracle as O  # noqa: E402


def check(name, n, quick=False):
    prog = oracle_prog(name)
    t0 = time.time()
    oxyz, oidx, oact = O.extract_sdf(n, prog)
    t_or = time.time() - t0
    mc = iso.MarchingCubes(n)
    t0 = time.time()
    nv, nt, na = mc.extract_device(iso.Sampler(iso_source(name)))
    t_gpu = time.time() - t0
    xyz, idx = mc.copy_out()
    msg = mesh_diff(xyz, idx, oxyz, oidx)
    ok = (msg == "" and na == oact)
    print("%-16s N=%-4d sdf  V=%d T=%d act=%d (oracle %d %d %d)  %s  [oracle %.2fs gpu %.3fs]"
          % (name, n, nv, nt, na, len(oxyz) // 3, len(oidx) // 3, oact, "OK" if ok else "MISMATCH: " + msg, t_or, t_gpu), flush=True)
    if not ok and not quick:
        grid = O.fill_grid_sdf(n, prog)
        ci_o = O.cube_indices(n, grid)
        ci_g = mc.cube_indices()
        bad = np.argwhere(ci_o != ci_g)
        print("   cube_index mismatches: %d of %d; first %s" % (len(bad), ci_o.size, bad[:5].tolist()))
    # grid-backed with the same samples
    grid = O.fill_grid_sdf(n, prog)
    mc2 = iso.MarchingCubes(n)
    nv2, nt2, na2 = mc2.extract_device(iso.DenseGrid(grid))
    xyz2, idx2 = mc2.copy_out()
    msg2 = mesh_diff(xyz2, idx2, oxyz, oidx)
    print("%-16s N=%-4d grid V=%d T=%d act=%d  %s" % (name, n, nv2, nt2, na2, "OK" if msg2 == "" and na2 == oact else "MISMATCH: " + msg2), flush=True)
    mc.close(); mc2.close()
    return ok and msg2 == ""


if __name__ == "__main__":
    allok = True
    for name, n in [("sphere03", 8), ("sphere03", 32), ("sphere05_origin", 32), ("torus", 64), ("csgA", 64), ("csgB", 64),
                    ("torus_origin", 128), ("nested", 100), ("prism", 33), ("cylinder", 65), ("sphere03", 2), ("sphere03", 3),
                    ("torus", 256), ("csgA", 256)]:
        try:
            allok &= check(name, n)
        except Exception as e:  # keep going: first contact wants the whole picture
            allok = False
            print("%-16s N=%-4d EXCEPTION %r" % (name, n, e), flush=True)
    print("ALL OK" if allok else "FAILURES")
    sys.exit(0 if allok else 1)
