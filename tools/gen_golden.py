#!/usr/bin/env python3
"""Generates tests/golden/mesh_hashes.json with the CPU oracle (the Rust reference cannot run in this
image; the oracle is pinned to the SURVEY.md 8(c) known answers, which this file also records)."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import oracle_prog, sha  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [("torus_origin", 128), ("sphere03", 32), ("sphere05_origin", 32), ("sphere03", 64), ("torus", 64),
         ("csgA", 64), ("csgB", 64), ("sphere03", 256), ("torus", 256), ("csgA", 256), ("csgB", 256),
         # beyond the survey's list (oracle-generated): non-power-of-two sizes, other shapes
         ("prism", 33), ("cylinder", 65), ("nested", 100), ("csgA", 97), ("torus", 200), ("sphere03", 3)]

out = []
for name, n in CASES:
    xyz, idx, act = O.extract_sdf(n, oracle_prog(name), O.LEAN)
    out.append({"shape": name, "size": n, "active_cells": act, "vertices": len(xyz) // 3, "triangles": len(idx) // 3,
                "sha_v": sha(xyz, "<f4"), "sha_i": sha(idx, "<u4")})
    print(out[-1])
(ROOT / "tests" / "golden" / "mesh_hashes.json").write_text(json.dumps(out, indent=1) + "\n")
