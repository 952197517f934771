"""The CPU oracle against everything that pins it: the reference's own SDF known-answer tests, the
SURVEY.md 8(c) known answers (independent restatement), committed golden hashes and mesh invariants."""
import json
from pathlib import Path

import numpy as np
import pytest

from helpers import mesh_invariants, oracle_prog, sha

GOLDEN = json.loads((Path(__file__).parent / "golden" / "mesh_hashes.json").read_text())

# (shape, N, active, V, T, sha_v prefix, sha_i prefix)
SURVEY_PINS = [
    ("torus_origin", 128, 2859, 2974, 5718, "e1806e706f30", "702f461a7632"),
    ("sphere03", 32, 1610, 1608, 3212, "9e60fdff8f2d", "1fc53132926f"),
    ("sphere05_origin", 32, 562, 609, 1123, "ae5bf2b9d16a", "f17a5cc84e8e"),
    ("sphere03", 64, 6746, 6744, 13484, "8d4803d188361974", "55d803561a4d9057"),
    ("torus", 64, 5528, 5528, 11056, "dbe15c91f8a29175", "515721d85539a43c"),
    ("csgA", 64, 4818, 4824, 9660, "bb93e5ee238d6b2f", "3b6ac087e05cdfba"),
    ("csgB", 64, 4058, 4056, 8108, "ad2f6e21d6949477", "90c691d3eeaf3256"),
    ("sphere03", 256, 110162, 110160, 220316, "b6c771ba671b5b19", "f38da6363cb8a759"),
    ("torus", 256, 92032, 92032, 184064, "cedd63ac2dcaa506", "fad728428df72163"),
    ("csgA", 256, 71370, 71376, 142764, "8217da29339884bf", "b651b71d3b77332f"),
    ("csgB", 256, 62426, 62424, 124844, "c15197f25ba2f530", "eec61632d8a56b2a"),
]


def test_reference_sdf_known_answers(oracle):
    """exact-equality asserts of the reference's unit tests (scalar side)"""
    O = oracle
    MAXF = np.float32(3.4028235e38)

    def s(prog, p):
        return float(O.sample_sdf(O.program(prog), np.array([p], np.float32))[0])

    sph = [(O.SPHERE, 2.0)]  # sphere.rs:70-76
    assert s(sph, (0, 0, 0)) == -2.0 and s(sph, (2, 0, 0)) == 0.0 and s(sph, (0, 0, 8)) == 6.0 and s(sph, (8, 0, 0)) == 6.0
    tor = [(O.TORUS, 8.0, 2.0)]  # torus.rs:115-122
    assert s(tor, (0, 0, 0)) == 6.0 and s(tor, (8, 0, 0)) == -2.0 and s(tor, (10, 0, 0)) == 0.0
    assert s(tor, (12, 0, 0)) == 2.0 and s(tor, (8, 0, 8)) == 6.0
    cyl = [(O.CYLINDER, 2.0, 4.0)]  # cylinder.rs:93-99
    assert s(cyl, (0, 0, 0)) == -2.0 and s(cyl, (2, 0, 4)) == 0.0 and s(cyl, (0, 0, 8)) == 4.0 and s(cyl, (8, 0, 0)) == 6.0
    pr = [(O.PRISM, 1.0, 2.0, 4.0)]  # rectangular_prism.rs:87-93
    assert s(pr, (0, 0, 0)) == -1.0 and s(pr, (1, 2, 4)) == 0.0 and s(pr, (0, 0, 8)) == 4.0 and s(pr, (8, 0, 0)) == 7.0
    a, b = (O.PRISM, 4.0, 4.0, 1.0), (O.PRISM, 2.0, 2.0, 4.0)  # csg.rs:115-155
    u = [a, b, (O.UNION,)]
    assert s(u, (0, 0, 0)) == -2.0 and s(u, (4, 4, 1)) == 0.0 and s(u, (0, 0, 8)) == 4.0 and s(u, (8, 0, 0)) == 4.0
    i = [a, b, (O.INTERSECTION,)]
    assert s(i, (0, 0, 0)) == -1.0 and s(i, (2, 2, 1)) == 0.0 and s(i, (0, 0, 8)) == 7.0 and s(i, (8, 0, 0)) == 6.0
    d = [a, b, (O.DIFFERENCE,)]  # csg.rs:96-100: max(b, -a)
    assert s(d, (0, 0, 0)) == max(-2.0, 1.0) and s(d, (0, 0, 3)) == max(-1.0, -2.0)
    t = [(O.TRANSLATE_PUSH, .5, .5, .5), (O.SPHERE, 2.0), (O.TRANSLATE_POP,)]  # examples/common/sources.rs:38-43
    assert s(t, (.5, .5, .5)) == -2.0 and s(t, (2.5, .5, .5)) == 0.0
    del MAXF


@pytest.mark.parametrize("shape,n,act,V,T,sv,si", SURVEY_PINS, ids=lambda v: str(v)[:14])
def test_survey_known_answers(oracle, shape, n, act, V, T, sv, si):
    xyz, idx, a = oracle.extract_sdf(n, oracle_prog(shape), oracle.LEAN)
    assert (a, len(xyz) // 3, len(idx) // 3) == (act, V, T)
    assert sha(xyz, "<f4").startswith(sv) and sha(idx, "<u4").startswith(si)


@pytest.mark.parametrize("g", [g for g in GOLDEN if g["size"] <= 128], ids=lambda g: "%s-%d" % (g["shape"], g["size"]))
def test_faithful_mode_equals_lean_and_golden(oracle, g):
    """the faithful-cost mode (SipHash'd 48-byte keys + edge bookkeeping) and the lean mode are the same algorithm"""
    for mode in (oracle.FAITHFUL, oracle.LEAN):
        xyz, idx, act = oracle.extract_sdf(g["size"], oracle_prog(g["shape"]), mode)
        assert act == g["active_cells"] and len(xyz) // 3 == g["vertices"] and len(idx) // 3 == g["triangles"]
        assert sha(xyz, "<f4") == g["sha_v"] and sha(idx, "<u4") == g["sha_i"]


def test_grid_source_equals_sdf_source(oracle):
    for shape, n in (("csgA", 48), ("torus", 37)):
        prog = oracle_prog(shape)
        xyz, idx, act = oracle.extract_sdf(n, prog)
        grid = oracle.fill_grid_sdf(n, prog)
        assert grid.shape == (n + 1, n, n)
        gx, gi, ga = oracle.extract_grid(n, grid)
        assert np.array_equal(gi, idx) and np.array_equal(gx.view(np.uint32), xyz.view(np.uint32)) and ga == act


def test_z_window_is_a_prefix(oracle):
    """extract_grid with z_cells < size (the bounded CPU-baseline sample) is the exact prefix in cell order"""
    n = 40
    grid = oracle.fill_grid_sdf(n, oracle_prog("torus"))
    xyz, idx, _ = oracle.extract_grid(n, grid)
    wx, wi, _ = oracle.extract_grid(n, grid[:21], z_cells=20)
    assert np.array_equal(wi, idx[:len(wi)]) and np.array_equal(wx, xyz[:len(wx)])
    assert 0 < len(wi) < len(idx)


def test_invariants_closed_surfaces(oracle):
    # sphere: T = 2V - 4 (Euler characteristic 2); torus: T = 2V (0).  Centred shapes stay inside the
    # lattice incl. the extra z layer, so the meshes are closed and consistently oriented.
    for shape, n, euler in (("sphere03", 48, 2), ("torus", 64, 0), ("csgB", 50, 2)):
        xyz, idx, _ = oracle.extract_sdf(n, oracle_prog(shape))
        f = mesh_invariants(xyz, idx, closed=True)
        assert f["euler"] == euler and f["unpaired"] == 0 and f["directed_edge_dups"] == 0, f
        assert f["T"] == 2 * f["V"] - 2 * euler


def test_z_quirk_extra_layer(oracle):
    """primal_grid.rs:59 runs z over 0..size: Sphere(0.6)@0.5, N=32 has 176 of 3676 active cells in layer z = N-1"""
    n = 32
    prog = oracle.program([(oracle.TRANSLATE_PUSH, .5, .5, .5), (oracle.SPHERE, .6), (oracle.TRANSLATE_POP,)])
    grid = oracle.fill_grid_sdf(n, prog)
    ci = oracle.cube_indices(n, grid)
    active = (ci != 0) & (ci != 255)
    assert ci.shape == (n, n - 1, n - 1)
    assert int(active.sum()) == 3676 and int(active[n - 1].sum()) == 176


def test_edge_cases(oracle):
    O = oracle
    # size 1: no cells; size 2: one cell column of 2 cells
    xyz, idx, act = O.extract_sdf(1, oracle_prog("sphere03"))
    assert len(xyz) == 0 and len(idx) == 0 and act == 0
    # all-outside / all-inside fields are empty
    for v in (1.0, -1.0):
        g = np.full((9, 8, 8), v, np.float32)
        xyz, idx, act = O.extract_grid(8, g)
        assert len(xyz) == 0 and len(idx) == 0 and act == 0
    # +0.0, -0.0 and NaN all count as inside (`!(v > 0)`, marching_cubes_impl.rs:32)
    for special in (0.0, -0.0, np.nan):
        g = np.full((5, 4, 4), 1.0, np.float32)
        g[2, 2, 2] = special
        ci = O.cube_indices(4, g)
        assert int(((ci != 0) & (ci != 255)).sum()) == 8
    # +inf is outside, -inf inside; infinities give NaN/edge positions but a well-formed topology
    g = np.full((5, 4, 4), np.inf, np.float32)
    g[2, 2, 2] = -np.inf
    xyz, idx, act = O.extract_grid(4, g)
    assert act == 8 and len(idx) // 3 == 8 and len(xyz) // 3 == 6


def test_point_cloud_oracle_properties(oracle):
    """PointCloud restatement (reference src/point_cloud.rs:50-63): one point per active cell, cell order, cell centres"""
    from helpers import oracle_prog
    for name, size in (("sphere03", 32), ("torus", 40)):
        prog = oracle_prog(name)
        pts = oracle.point_cloud_sdf(size, prog).reshape(-1, 3)
        _, _, act = oracle.extract_sdf(size, prog)
        assert len(pts) == act                                   # same active-cell predicate as MarchingCubes
        grid = oracle.fill_grid_sdf(size, prog)
        assert oracle.point_cloud_grid(size, grid).tobytes() == pts.tobytes()
        ci = oracle.cube_indices(size, grid)
        z, y, x = np.nonzero((ci != 0) & (ci != 255))            # C order == (z, y, x) cell order
        inv = np.float32(1.0) / np.float32(size - 1)
        half = np.float32(0.5)
        want = np.stack([half * (c.astype(np.float32) * inv) + half * ((c + 1).astype(np.float32) * inv) for c in (x, y, z)], axis=1)
        assert want.astype(np.float32).tobytes() == pts.tobytes()
    assert oracle.point_cloud_sdf(1, oracle_prog("sphere03")).size == 0


def test_central_difference_normals_oracle(oracle):
    """IndexedInterleavedNormals over CentralDifference (reference src/extractor.rs:113-122, src/source.rs:82-94):
    float64 central differences agree to the precision f32 differences at epsilon = 1e-6 allow; layout x y z nx ny nz"""
    inner = oracle.program([(oracle.SPHERE, .3)])
    xyz, _, _ = oracle.extract_sdf(24, oracle.program([(oracle.TRANSLATE_PUSH, .5, .5, .5), (oracle.SPHERE, .3), (oracle.TRANSLATE_POP,)]))
    out = oracle.interleaved_normals_cd(inner, xyz, 0.000001, [(.5, .5, .5)])
    assert out.shape == (len(xyz) // 3, 6) and out[:, :3].tobytes() == xyz.tobytes()
    q = xyz.reshape(-1, 3).astype(np.float64) - 0.5
    want = q / np.linalg.norm(q, axis=1, keepdims=True)          # gradient of |q| - r
    assert np.abs(out[:, 3:] - want).max() < 0.05                # f32 differences over 2e-6: a few ulps of 0.3 / 2e-6
    # a larger epsilon is accurate
    out2 = oracle.interleaved_normals_cd(inner, xyz, 0.001, [(.5, .5, .5)])
    assert np.abs(out2[:, 3:] - want).max() < 1e-3


def test_directed_sample_vector_known_answers(oracle):
    """the reference's exact-equality unit tests of VectorSource::sample_vector (sphere.rs:78-93, torus.rs:124-155,
    cylinder.rs:101-124, rectangular_prism.rs:95-115, csg.rs:126-136,157-167) pin the Directed restatement"""
    O = oracle
    MAX = np.finfo(np.float32).max
    f = np.float32

    def sv(nodes, p):
        return tuple(O.sample_sdf_vector(O.program(nodes), [p])[0].tolist())

    sph = [(O.SPHERE, 2.0)]
    assert sv(sph, (0, 0, 0)) == (-2.0, -2.0, -2.0)
    assert sv(sph, (2, 0, 0)) == (0.0, 0.0, 0.0)
    assert sv(sph, (0, 0, 8)) == (MAX, MAX, 6.0)
    assert sv(sph, (8, 0, 0)) == (6.0, MAX, MAX)
    tor = [(O.TORUS, 8.0, 2.0)]
    assert sv(tor, (0, 0, 0)) == (6.0, 6.0, MAX)
    assert sv(tor, (8, 0, 0)) == (-2.0, -2.0, -2.0)
    assert sv(tor, (10, 0, 0)) == (0.0, 0.0, 0.0)
    assert sv(tor, (12, 0, 0)) == (2.0, MAX, MAX)
    assert sv(tor, (12, 12, 0)) == (MAX, MAX, MAX)
    v = float(f(9.0) - np.sqrt(f(10.0 * 10.0 - 9.0 * 9.0)))
    assert sv(tor, (9, 9, 0)) == (v, v, MAX)
    v = float(np.sqrt(f(6.0 * 6.0 - 2.0 * 2.0)) - f(2.0))
    assert sv(tor, (2, 2, 0)) == (v, v, MAX)
    assert sv(tor, (8, 0, 8)) == (MAX, MAX, 6.0)
    cyl = [(O.CYLINDER, 2.0, 4.0)]
    assert sv(cyl, (0, 0, 0)) == (-2.0, -2.0, -4.0)
    assert sv(cyl, (0, 0, 1)) == (-2.0, -2.0, -3.0)
    v = float(f(1.0) - np.sqrt(f(4.0) - f(1.0)))
    assert sv(cyl, (1, 1, 1)) == (v, v, -3.0)
    assert sv(cyl, (2, 0, 4)) == (0.0, 0.0, 0.0)
    assert sv(cyl, (0, 0, 8)) == (MAX, MAX, 4.0)
    assert sv(cyl, (8, 0, 0)) == (6.0, MAX, MAX)
    pri = [(O.PRISM, 1.0, 2.0, 4.0)]
    assert sv(pri, (0, 0, 0)) == (-1.0, -2.0, -4.0)
    assert sv(pri, (1, 2, 4)) == (0.0, 0.0, 0.0)
    assert sv(pri, (0, 0, 8)) == (MAX, MAX, 4.0)
    assert sv(pri, (8, 0, 0)) == (7.0, MAX, MAX)
    assert sv(pri, (8, 8, 8)) == (MAX, MAX, MAX)
    a, b = (O.PRISM, 4.0, 4.0, 1.0), (O.PRISM, 2.0, 2.0, 4.0)
    uni = [a, b, (O.UNION,)]
    assert sv(uni, (0, 0, 0)) == (-4.0, -4.0, -4.0)
    assert sv(uni, (4, 4, 1)) == (0.0, 0.0, 0.0)
    assert sv(uni, (0, 0, 8)) == (MAX, MAX, 4.0)
    assert sv(uni, (8, 0, 0)) == (4.0, MAX, MAX)
    inter = [a, b, (O.INTERSECTION,)]
    assert sv(inter, (0, 0, 0)) == (-2.0, -2.0, -1.0)
    assert sv(inter, (2, 2, 1)) == (0.0, 0.0, 0.0)
    assert sv(inter, (0, 0, 8)) == (MAX, MAX, 7.0)
    assert sv(inter, (8, 0, 0)) == (6.0, MAX, MAX)


def test_directed_extract_oracle_sanity(oracle):
    """MarchingCubes<Directed> restatement: a closed surface near the Signed one (same shape, different distance field)"""
    from helpers import oracle_prog, mesh_invariants
    for name in ("sphere03", "torus", "csgB"):
        xyz, idx, act = oracle.extract_sdf_directed(32, oracle_prog(name))
        sxyz, sidx, sact = oracle.extract_sdf(32, oracle_prog(name))
        assert len(idx) > 0 and act > 0
        facts = mesh_invariants(xyz, idx, closed=False)
        assert facts["directed_edge_dups"] == 0
        assert abs(len(idx) - len(sidx)) < 0.2 * len(sidx)          # same surface, slightly different cell set at most


EXTRAS = json.loads((Path(__file__).parent / "golden" / "extras_hashes.json").read_text())


def test_widened_rows_match_committed_golden_hashes(oracle):
    """PointCloud, MarchingCubes<Directed> and interleaved central-difference normals against tests/golden/extras_hashes.json
    (tools/gen_golden_extras.py): the restatements may not drift silently"""
    from helpers import oracle_prog, sha
    for g in EXTRAS["point_cloud"]:
        if g["size"] <= 100:
            pts = oracle.point_cloud_sdf(g["size"], oracle_prog(g["shape"]))
            assert (len(pts) // 3, sha(pts, "<f4")) == (g["points"], g["sha_p"]), g
    for g in EXTRAS["directed"]:
        if g["size"] <= 65:
            xyz, idx, act = oracle.extract_sdf_directed(g["size"], oracle_prog(g["shape"]))
            assert (act, len(xyz) // 3, len(idx) // 3, sha(xyz, "<f4"), sha(idx, "<u4")) == \
                (g["active_cells"], g["vertices"], g["triangles"], g["sha_v"], g["sha_i"]), g
    O = oracle
    inner = {"torus": [(O.TORUS, .25, .1)],
             "csgA": [(O.SPHERE, .25), (O.PRISM, .2, .2, .2), (O.DIFFERENCE,), (O.CYLINDER, .02, .25), (O.UNION,)],
             "csgB": [(O.SPHERE, .3), (O.PRISM, .2, .2, .2), (O.INTERSECTION,)]}
    for g in EXTRAS["normals"]:
        xyz, _, _ = O.extract_sdf(g["size"], oracle_prog(g["shape"]))
        xyzn = O.interleaved_normals_cd(O.program(inner[g["shape"]]), xyz, g["epsilon"], [(.5, .5, .5)])
        assert (len(xyzn), sha(xyzn, "<f4")) == (g["vertices"], g["sha_vn"]), g


def test_directed_point_cloud_restatement(oracle):
    """PointCloud<Directed> (point_cloud.rs:50-63 with distance.rs:77-80): one point per cell whose corners are not all on one side
    under "outside iff any component of the directed distance is positive"; cross-checked against the vertex-sharing Directed
    extract (same active-cell count) and against a numpy formulation from sample_vector"""
    from helpers import oracle_prog
    for name, size in (("sphere03", 20), ("csgA", 24), ("torus", 22)):
        prog = oracle_prog(name)
        pts = oracle.point_cloud_sdf_directed(size, prog)
        _, _, act = oracle.extract_sdf_directed(size, prog)
        assert len(pts) // 3 == act and act > 0
        inv = np.float32(1.0) / np.float32(size - 1)
        ax = (np.arange(size, dtype=np.float32) * inv).astype(np.float32)
        zs = (np.arange(size + 1, dtype=np.float32) * inv).astype(np.float32)
        Z, Y, X = np.meshgrid(zs, ax, ax, indexing="ij")
        P = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1).astype(np.float32)
        V = oracle.sample_sdf_vector(prog, P).reshape(size + 1, size, size, 3)
        inside = ~((V[..., 0] > 0) | (V[..., 1] > 0) | (V[..., 2] > 0))
        n_in = sum(inside[dz:size + dz, dy:size - 1 + dy, dx:size - 1 + dx].astype(np.int32) for dz in (0, 1) for dy in (0, 1) for dx in (0, 1))
        active = (n_in != 0) & (n_in != 8)
        zz, yy, xx = np.nonzero(active)
        half = np.float32(0.5)
        want = np.stack([half * ax[xx] + half * ax[xx + 1], half * ax[yy] + half * ax[yy + 1], half * zs[zz] + half * zs[zz + 1]], axis=1)
        assert np.array_equal(pts.reshape(-1, 3).view(np.uint32), want.astype(np.float32).view(np.uint32)), name

