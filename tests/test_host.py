"""Host-side logic (no GPU): source encoding, extractor protocol, slab partition and base arithmetic."""
import numpy as np
import pytest

import isosurface_b200 as iso
from isosurface_b200 import _lib
from isosurface_b200.extractor import replay
from isosurface_b200.sharded import bases_from_totals, slab_range, slab_sample_layers
from isosurface_b200.source import encode_program


def test_program_encoding_matches_oracle_programs(oracle):
    from helpers import iso_source, oracle_prog
    for name in ("torus_origin", "sphere03", "torus", "csgA", "csgB", "prism", "cylinder", "nested"):
        got = encode_program(iso.Sampler(iso_source(name)))
        want = oracle_prog(name)
        assert got.dtype == want.dtype and got.tobytes() == want.tobytes(), name


def test_difference_operand_order():
    # Difference{a, b} = max(b, -a): a is pushed first (csg.rs:96-100)
    prog = encode_program(iso.Difference(iso.Sphere(1.0), iso.Torus(2.0, 0.5)))
    assert [int(n["op"]) for n in prog] == [_lib.SDF_SPHERE, _lib.SDF_TORUS, _lib.SDF_DIFFERENCE]


def test_arbitrary_callables_are_rejected():
    with pytest.raises(TypeError, match="not a device source"):
        encode_program(iso.Sampler(lambda p: p[0]))
    with pytest.raises(TypeError):
        iso.Union(iso.Sphere(1.0), object())


def test_dense_grid_validation():
    g = np.zeros((9, 8, 8), np.float32)
    d = iso.DenseGrid(g)
    assert d.size == 8 and not d.on_device
    with pytest.raises(ValueError):
        iso.DenseGrid(np.zeros((8, 8, 8), np.float32))     # needs N+1 layers (primal_grid.rs:59)
    with pytest.raises(ValueError):
        iso.DenseGrid(1234, size=None, on_device=True)


class Recorder(iso.Extractor):
    def __init__(self):
        self.calls = []

    def extract_vertex(self, v):
        self.calls.append(("v", v))

    def extract_index(self, i):
        self.calls.append(("i", i))


def test_extractor_protocol_order():
    """all vertices first, in order, then all indices (marching_cubes.rs:81, mesh.rs:91-100)"""
    xyz = np.arange(12, dtype=np.float32)
    idx = np.array([0, 1, 2, 2, 1, 3], np.uint32)
    r = Recorder()
    replay(r, xyz, idx)
    kinds = [c[0] for c in r.calls]
    assert kinds == ["v"] * 4 + ["i"] * 6
    assert r.calls[1][1] == (3.0, 4.0, 5.0) and [c[1] for c in r.calls[4:]] == idx.tolist()
    vs, is_ = [], []
    replay(iso.IndexedVertices(vs, is_), xyz, idx)
    assert vs == xyz.tolist() and is_ == idx.tolist()
    only = []
    replay(iso.OnlyVertices(only), xyz, idx)
    assert only == xyz.tolist()


def test_slab_ranges_partition_the_layers():
    for size in (2, 7, 64, 513, 2048):
        for world in (1, 2, 3, 4, 8):
            if world > size:
                continue
            rs = [slab_range(size, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == size
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            lens = [b - a for a, b in rs]
            assert max(lens) - min(lens) <= 1 and min(lens) >= 1
            for r in range(world):
                z_first, n = slab_sample_layers(size, r, world)
                assert z_first == rs[r][0] - (1 if r else 0) and z_first + n - 1 == rs[r][1]


def test_bases_from_totals():
    g = np.array([[10, 7, 20], [5, 1, 9], [8, 8, 16]], dtype=np.uint64)
    assert bases_from_totals(g, 0) == (0, 0, 0)
    assert bases_from_totals(g, 1) == (10, 7, 20)          # boundary = first vertex of rank 0's last layer
    assert bases_from_totals(g, 2) == (15, 11, 29)


def test_new_host_side_argument_checks_need_no_gpu():
    """argument validation of the widened API happens before any device call"""
    import isosurface_b200 as iso
    from isosurface_b200.source import find_central_difference
    cd = iso.CentralDifference(iso.Sphere(0.3), 0.001)
    assert find_central_difference(iso.Sampler(iso.Translate(0.5, cd))) is cd
    assert find_central_difference(iso.Sampler(iso.Translate(0.5, iso.Sphere(0.3)))) is None
    assert find_central_difference(iso.Union(cd, iso.Sphere(0.1))) is None      # only enclosing wrappers are looked through
    with pytest.raises(TypeError):
        iso.IndexedInterleavedNormals([], [], iso.Sampler(iso.Sphere(0.3)))
    sink = iso.IndexedInterleavedNormals([], [], iso.Sampler(cd))
    assert sink.central_difference.epsilon == 0.001
    with pytest.raises(ValueError):
        iso.MarchingCubes(32, distance="euclidean")
    with pytest.raises(ValueError):
        iso.ChunkedMarchingCubes(32, n_inflight=0)
    # CentralDifference is transparent as a scalar source: same program as the tree it wraps
    from isosurface_b200.source import encode_program
    assert encode_program(iso.Translate(0.5, cd)).tobytes() == encode_program(iso.Translate(0.5, iso.Sphere(0.3))).tobytes()


def test_balanced_slabs_cut_equal_work():
    """sharded.balanced_slabs: contiguous, covering, at least one layer per rank, and the heaviest rank's work is close to the mean
    (never worse than the equal-thickness split on a skewed field)"""
    from isosurface_b200.sharded import balanced_slabs, slab_range
    size = 256
    rng = np.random.default_rng(3)
    z = np.arange(size)
    active = (40000 * np.exp(-((z - 70) / 25.0) ** 2) + 3000 * rng.random(size)).astype(np.int64)  # the surface sits in one region
    for world in (1, 2, 3, 8, 31):
        slabs = balanced_slabs(size, active, world)
        assert slabs[0][0] == 0 and slabs[-1][1] == size and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
        assert all(z1 > z0 for z0, z1 in slabs)
        work = size * size + 76.0 * active
        heavy = max(work[z0:z1].sum() for z0, z1 in slabs)
        heavy_equal = max(work[slice(*slab_range(size, r, world))].sum() for r in range(world))
        assert heavy <= heavy_equal + work.max()
        if world in (2, 3, 8):
            assert heavy <= 1.15 * work.sum() / world
    # a uniform field keeps the equal split; degenerate inputs are rejected
    assert balanced_slabs(64, np.full(64, 100), 4) == [slab_range(64, r, 4) for r in range(4)]
    assert balanced_slabs(5, np.zeros(5), 5) == [(i, i + 1) for i in range(5)]
    with pytest.raises(ValueError):
        balanced_slabs(8, np.zeros(7), 2)
    with pytest.raises(ValueError):
        balanced_slabs(4, np.zeros(4), 5)
