#!/usr/bin/env python3
"""Generates tests/golden/extras_hashes.json with the CPU oracle: PointCloud, MarchingCubes<Directed> and
IndexedInterleavedNormals over CentralDifference (the Rust reference cannot run in this image; see the pinning notes in
oracle/mc_oracle.c).  SHA-256 over the little-endian f32 / u32 streams in emission order."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import oracle_prog, sha  # noqa: E402
from oracle import oracle as O  # noqa: E402

INNER = {  # the trees inside the (0.5, 0.5, 0.5) translation, for the normals
    "torus": [(O.TORUS, .25, .1)],
    "csgA": [(O.SPHERE, .25), (O.PRISM, .2, .2, .2), (O.DIFFERENCE,), (O.CYLINDER, .02, .25), (O.UNION,)],
    "csgB": [(O.SPHERE, .3), (O.PRISM, .2, .2, .2), (O.INTERSECTION,)],
}
out = {"point_cloud": [], "directed": [], "normals": []}
for name, n in [("sphere03", 32), ("torus_origin", 128), ("csgA", 64), ("csgB", 64), ("nested", 100)]:
    pts = O.point_cloud_sdf(n, oracle_prog(name))
    out["point_cloud"].append({"shape": name, "size": n, "points": len(pts) // 3, "sha_p": sha(pts, "<f4")})
for name, n in [("sphere03", 32), ("torus", 64), ("csgA", 64), ("csgB", 64), ("cylinder", 65), ("prism", 33), ("torus_origin", 128)]:
    xyz, idx, act = O.extract_sdf_directed(n, oracle_prog(name))
    out["directed"].append({"shape": name, "size": n, "active_cells": act, "vertices": len(xyz) // 3, "triangles": len(idx) // 3,
                            "sha_v": sha(xyz, "<f4"), "sha_i": sha(idx, "<u4")})
for name, n, eps in [("torus", 64, 0.000001), ("csgA", 64, 0.000001), ("csgB", 64, 0.001)]:
    xyz, idx, _ = O.extract_sdf(n, oracle_prog(name))
    xyzn = O.interleaved_normals_cd(O.program(INNER[name]), xyz, eps, [(.5, .5, .5)])
    out["normals"].append({"shape": name, "size": n, "epsilon": eps, "vertices": len(xyzn), "sha_vn": sha(xyzn, "<f4")})
for k, v in out.items():
    for r in v:
        print(k, r)
(ROOT / "tests" / "golden" / "extras_hashes.json").write_text(json.dumps(out, indent=1) + "\n")
