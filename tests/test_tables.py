"""Table integrity (SURVEY.md 8c pins) for the oracle's tables and the product's derived tables."""
import hashlib
import json
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

PINS = {"CORNERS": "fc803b129ace5681", "EDGE_CONNECTION": "75a37b753a8ad6ae",
        "EDGE_CROSSING_MASK": "18bd4425530b89a2", "TRIANGLE_CONNECTION": "57353f8cdd905651"}


def _sha(table):
    return hashlib.sha256(json.dumps(table, separators=(",", ":")).encode()).hexdigest()[:16]


def test_oracle_tables_match_survey_pins(oracle):
    tri, em, co, ed = oracle.tables()
    assert _sha(co.tolist()) == PINS["CORNERS"]
    assert _sha(ed.tolist()) == PINS["EDGE_CONNECTION"]
    assert _sha([int(x) for x in em]) == PINS["EDGE_CROSSING_MASK"]
    assert _sha(tri.tolist()) == PINS["TRIANGLE_CONNECTION"]


def test_oracle_table_structure(oracle):
    tri, em, _, _ = oracle.tables()
    ntri = [(row >= 0).sum() // 3 for row in tri]
    assert sum(ntri) == 820
    hist = {k: ntri.count(k) for k in range(6)}
    assert hist == {0: 2, 1: 16, 2: 50, 3: 80, 4: 76, 5: 32}
    for c in range(256):
        assert int(em[c]) == sum(1 << e for e in set(int(v) for v in tri[c] if v >= 0))


def _product_tables(tmp_path):
    exe = tmp_path / "dump_tables"
    subprocess.run(["g++", "-O1", "-o", str(exe), str(ROOT / "tests" / "dump_tables.cpp")], check=True)
    raw = subprocess.run([str(exe)], check=True, capture_output=True).stdout
    dt = np.dtype([("tri", "<u8", 256), ("order", "<u8", 256), ("before", "<u2", (256, 12)), ("emask", "<u2", 256),
                   ("ntri", "u1", 256), ("ownmask", "<u2", 8), ("owner", "u1", (8, 12)), ("ends", "u1", 12),
                   ("ref_of_nat", "u1", 256), ("rank3", "u1", 256), ("pad", "u1", 4)])
    assert len(raw) == dt.itemsize, (len(raw), dt.itemsize)
    return np.frombuffer(raw, dtype=dt)[0]


def test_product_tables_agree_with_oracle(oracle, tmp_path):
    t = _product_tables(tmp_path)
    tri, em, co, ed = oracle.tables()
    nat_of_ref = [int(c[0]) | int(c[1]) << 1 | int(c[2]) << 2 for c in co]
    for cref in range(256):
        cnat = sum(1 << nat_of_ref[i] for i in range(8) if cref >> i & 1)
        assert t["ref_of_nat"][cnat] == cref
        row = [int(v) for v in tri[cref] if v >= 0]
        packed = int(t["tri"][cnat])
        got = []
        while packed & 0xF != 0xF:
            got.append(packed & 0xF)
            packed >>= 4
        assert got == row
        assert t["ntri"][cnat] == len(row) // 3
        assert t["emask"][cnat] == em[cref]
        first = list(dict.fromkeys(row))
        order = [(int(t["order"][cnat]) >> (4 * k)) & 0xF for k in range(len(first))]
        assert order == first
        for k, e in enumerate(first):
            assert t["before"][cnat][e] == sum(1 << f for f in first[:k])
        own = [e for e in first if e in (5, 6, 10)]
        r3 = int(t["rank3"][cnat])
        for e, sh in ((5, 0), (6, 2), (10, 4)):
            if e in own:
                assert (r3 >> sh) & 3 == own.index(e)
    for e in range(12):
        u, v = co[ed[e][0]], co[ed[e][1]]
        assert t["ends"][e] == (int(u[0]) | int(u[1]) << 1 | int(u[2]) << 2) | (int(v[0]) | int(v[1]) << 1 | int(v[2]) << 2) << 4


def test_ownership_tables_match_first_containing_cell(oracle, tmp_path):
    """ownmask/owner must equal "the lexicographically first cell in (z,y,x) order containing the edge"
    (the creator under the reference's sequential HashMap dedup, index_cache.rs:49-60 + mesh.rs:240-251)."""
    t = _product_tables(tmp_path)
    _, _, co, ed = oracle.tables()
    n = 3  # cells per axis
    first = {}
    for z in range(n):
        for y in range(n):
            for x in range(n):
                for e in range(12):
                    a = tuple(int(v) for v in (np.array([x, y, z]) + co[ed[e][0]]))
                    b = tuple(int(v) for v in (np.array([x, y, z]) + co[ed[e][1]]))
                    key = (min(a, b), max(a, b))
                    first.setdefault(key, ((x, y, z), e))
    for z in range(n):
        for y in range(n):
            for x in range(n):
                flags = (x == 0) | (y == 0) << 1 | (z == 0) << 2
                for e in range(12):
                    a = tuple(int(v) for v in (np.array([x, y, z]) + co[ed[e][0]]))
                    b = tuple(int(v) for v in (np.array([x, y, z]) + co[ed[e][1]]))
                    (ox, oy, oz), oe = first[(min(a, b), max(a, b))]
                    ow = int(t["owner"][flags][e])
                    assert (x - (ow & 1), y - (ow >> 1 & 1), z - (ow >> 2 & 1)) == (ox, oy, oz)
                    assert ow >> 4 == oe
                    assert bool(t["ownmask"][flags] >> e & 1) == ((ox, oy, oz) == (x, y, z))


def test_generated_headers_are_current():
    """tools/gen_tables.py --check compares against the reference when it is mounted (not on the GPU box)."""
    if not Path("/root/reference/src/marching_cubes_tables.rs").exists():
        import pytest
        pytest.skip("reference tree not mounted")
    subprocess.run(["python", str(ROOT / "tools" / "gen_tables.py"), "--check"], check=True)
