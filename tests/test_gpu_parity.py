"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Bar: indices bit-exact (memcmp),
vertex positions bit-exact as well (the north star allows 1e-5 absolute; tolerance used: 0)."""
import ctypes as C
import json
import os
from pathlib import Path

import numpy as np
import pytest

from helpers import iso_source, mesh_diff, mesh_invariants, oracle_prog, sha

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((Path(__file__).parent / "golden" / "mesh_hashes.json").read_text())
POS_TOL = 0.0  # north star: 1e-5 absolute; achieved: identical bits


@pytest.fixture(scope="module")
def iso():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (the CUDA extension has no fallback)")
    import isosurface_b200
    return isosurface_b200


def synth(iso, kind, size, seed, z_first=0, n_layers=None):
    import torch
    from isosurface_b200 import _lib
    n_layers = size + 1 if n_layers is None else n_layers
    t = torch.empty(n_layers * size * size, dtype=torch.float32, device="cuda:0")
    _lib.check(_lib.load().isomc_synth_field(0, kind, size, seed, z_first, n_layers, C.c_void_p(t.data_ptr())))
    return t


def test_device_sdf_known_answers(iso, isolib):
    """the reference's exact-equality SDF unit tests, evaluated by the device interpreter"""
    from isosurface_b200 import _lib
    from isosurface_b200.source import encode_program

    def s(src, p):
        prog = encode_program(src)
        pts = np.array([p], np.float32)
        out = np.zeros(1, np.float32)
        _lib.check(isolib.isomc_debug_sample_sdf(0, prog.ctypes.data, len(prog), pts.ctypes.data, 1, out.ctypes.data))
        return float(out[0])

    sp = iso.Sphere(2.0)
    assert [s(sp, p) for p in ((0, 0, 0), (2, 0, 0), (0, 0, 8), (8, 0, 0))] == [-2.0, 0.0, 6.0, 6.0]
    to = iso.Torus(8.0, 2.0)
    assert [s(to, p) for p in ((0, 0, 0), (8, 0, 0), (10, 0, 0), (12, 0, 0), (8, 0, 8))] == [6.0, -2.0, 0.0, 2.0, 6.0]
    cy = iso.Cylinder(2.0, 4.0)
    assert [s(cy, p) for p in ((0, 0, 0), (2, 0, 4), (0, 0, 8), (8, 0, 0))] == [-2.0, 0.0, 4.0, 6.0]
    pr = iso.RectangularPrism((1.0, 2.0, 4.0))
    assert [s(pr, p) for p in ((0, 0, 0), (1, 2, 4), (0, 0, 8), (8, 0, 0))] == [-1.0, 0.0, 4.0, 7.0]
    a, b = iso.RectangularPrism((4.0, 4.0, 1.0)), iso.RectangularPrism((2.0, 2.0, 4.0))
    assert [s(iso.Union(a, b), p) for p in ((0, 0, 0), (4, 4, 1), (0, 0, 8), (8, 0, 0))] == [-2.0, 0.0, 4.0, 4.0]
    assert [s(iso.Intersection(a, b), p) for p in ((0, 0, 0), (2, 2, 1), (0, 0, 8), (8, 0, 0))] == [-1.0, 0.0, 7.0, 6.0]
    assert s(iso.Difference(a, b), (0, 0, 0)) == 1.0


def test_device_sdf_matches_oracle_bitwise(iso, isolib, oracle):
    from isosurface_b200 import _lib
    from isosurface_b200.source import encode_program
    rng = np.random.default_rng(7)
    pts = rng.uniform(-0.2, 1.2, size=(20000, 3)).astype(np.float32)
    for name in ("torus", "csgA", "csgB", "prism", "cylinder", "nested", "sphere05_origin"):
        prog = encode_program(iso_source(name))
        out = np.zeros(len(pts), np.float32)
        _lib.check(isolib.isomc_debug_sample_sdf(0, prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), out.ctypes.data))
        want = oracle.sample_sdf(oracle_prog(name), pts)
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32)), name
        # the stack-free chain evaluator the extract kernels use for left-deep trees: same bits
        out2 = np.zeros(len(pts), np.float32)
        _lib.check(isolib.isomc_debug_sample_sdf(0, prog.ctypes.data, len(prog) | 0x80000000, pts.ctypes.data, len(pts), out2.ctypes.data))
        assert np.array_equal(out2.view(np.uint32), want.view(np.uint32)), name + " (chain)"
    # a right-nested tree has no chain form: the generic interpreter runs it (and the chain request is refused)
    right = iso.Union(iso.Sphere(.2), iso.Intersection(iso.Translate(.5, iso.Sphere(.3)), iso.RectangularPrism((.2, .3, .4))))
    prog = encode_program(right)
    out = np.zeros(len(pts), np.float32)
    _lib.check(isolib.isomc_debug_sample_sdf(0, prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), out.ctypes.data))
    O = oracle
    oprog = O.program([(O.SPHERE, .2), (O.TRANSLATE_PUSH, .5, .5, .5), (O.SPHERE, .3), (O.TRANSLATE_POP,), (O.PRISM, .2, .3, .4),
                       (O.INTERSECTION,), (O.UNION,)])
    assert prog.tobytes() == oprog.tobytes()
    assert np.array_equal(out.view(np.uint32), O.sample_sdf(oprog, pts).view(np.uint32))
    assert isolib.isomc_debug_sample_sdf(0, prog.ctypes.data, len(prog) | 0x80000000, pts.ctypes.data, len(pts), out.ctypes.data) != 0
    mc = iso.MarchingCubes(48)
    mc.extract_device(iso.Sampler(right))
    xyz, idx = mc.copy_out()
    oxyz, oidx, _ = O.extract_sdf(48, oprog)
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    mc.close()


@pytest.mark.parametrize("g", GOLDEN, ids=lambda g: "%s-%d" % (g["shape"], g["size"]))
def test_sdf_extract_matches_golden_and_oracle(iso, oracle, g):
    """BASELINE configs C1a/C1b/C2 and more: committed hashes (incl. the SURVEY 8c pins) + live oracle"""
    mc = iso.MarchingCubes(g["size"])
    nv, nt, na = mc.extract_device(iso.Sampler(iso_source(g["shape"])))
    xyz, idx = mc.copy_out()
    assert (na, nv, nt) == (g["active_cells"], g["vertices"], g["triangles"])
    assert sha(idx, "<u4") == g["sha_i"], "index stream differs from golden"
    assert sha(xyz, "<f4") == g["sha_v"], "vertex stream differs from golden"
    if g["size"] <= 128:
        oxyz, oidx, oact = oracle.extract_sdf(g["size"], oracle_prog(g["shape"]))
        assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    mc.close()


def test_active_cell_sets_match(iso, oracle):
    """parity check #1: per-cell cube_index == oracle (active-cell sets bit-exact)"""
    for name, n in (("csgA", 70), ("torus_origin", 64)):
        grid = oracle.fill_grid_sdf(n, oracle_prog(name))
        mc = iso.MarchingCubes(n)
        mc.extract_device(iso.DenseGrid(grid))
        assert np.array_equal(mc.cube_indices(), oracle.cube_indices(n, grid))
        mc.close()


@pytest.mark.parametrize("kind,size,seed", [(1, 64, 0x1505F00D), (1, 129, 3), (2, 96, 0), (3, 100, 0x5EEDBA11), (1, 200, 11)],
                         ids=["fbm64", "fbm129", "gyroid96", "spheres100", "fbm200"])
def test_dense_grid_fields_match_oracle(iso, oracle, kind, size, seed):
    """grid-backed path on the synthetic bench fields (same bytes to both sides), incl. non-multiple-of-32 sizes"""
    t = synth(iso, kind, size, seed)
    host = t.cpu().numpy().reshape(size + 1, size, size)
    mc = iso.MarchingCubes(size)
    nv, nt, na = mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    oxyz, oidx, oact = oracle.extract_grid(size, host)
    assert na == oact
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    # host-buffer entry point gives the same bytes
    mc2 = iso.MarchingCubes(size)
    mc2.extract_device(iso.DenseGrid(host))
    x2, i2 = mc2.copy_out()
    assert np.array_equal(i2, idx) and np.array_equal(x2.view(np.uint32), xyz.view(np.uint32))
    mc.close(); mc2.close()


def test_edge_cases(iso, oracle):
    # sizes 1, 2, 3; all-inside / all-outside; special values
    for n in (1, 2, 3, 5):
        mc = iso.MarchingCubes(n)
        nv, nt, na = mc.extract_device(iso.Sampler(iso_source("sphere03")))
        oxyz, oidx, oact = oracle.extract_sdf(n, oracle_prog("sphere03"))
        xyz, idx = mc.copy_out()
        assert (nv, nt, na) == (len(oxyz) // 3, len(oidx) // 3, oact) and mesh_diff(xyz, idx, oxyz, oidx) == ""
        mc.close()
    for v in (1.0, -1.0, 0.0, -0.0, np.nan):
        g = np.full((34, 33, 33), v, np.float32)
        mc = iso.MarchingCubes(33)
        assert mc.extract_device(iso.DenseGrid(g)) == (0, 0, 0)
        mc.close()
    rng = np.random.default_rng(5)
    g = rng.standard_normal((41, 40, 40)).astype(np.float32)     # white noise: every case, max density
    g[rng.random(g.shape) < 0.05] = 0.0
    g[rng.random(g.shape) < 0.05] = -0.0
    g[rng.random(g.shape) < 0.02] = np.nan
    g[rng.random(g.shape) < 0.02] = np.inf
    g[rng.random(g.shape) < 0.02] = -np.inf
    mc = iso.MarchingCubes(40)
    nv, nt, na = mc.extract_device(iso.DenseGrid(g))
    xyz, idx = mc.copy_out()
    oxyz, oidx, oact = oracle.extract_grid(40, g)
    assert na == oact and np.array_equal(idx, oidx)
    # infinities/NaNs in the field give NaN coordinates; NaN *payloads* are hardware-specific (x86 SSE
    # produces 0xFFC00000, the GPU 0x7FFFFFFF), so: NaN where the oracle has NaN, identical bits elsewhere
    nan_g, nan_o = np.isnan(xyz), np.isnan(oxyz)
    assert np.array_equal(nan_g, nan_o) and nan_o.any()
    assert np.array_equal(xyz.view(np.uint32)[~nan_o], oxyz.view(np.uint32)[~nan_o])
    mc.close()


def test_dense_random_field_all_cases(iso, oracle):
    """white noise at a non-aligned size exercises all 256 cases, boundary ownership on every face and
    the worst-case compaction density"""
    rng = np.random.default_rng(11)
    n = 77
    g = rng.standard_normal((n + 1, n, n)).astype(np.float32)
    mc = iso.MarchingCubes(n)
    mc.extract_device(iso.DenseGrid(g))
    xyz, idx = mc.copy_out()
    oxyz, oidx, _ = oracle.extract_grid(n, g)
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    assert len(np.unique(oracle.cube_indices(n, g))) == 256
    mc.close()


def test_handle_reuse_and_determinism(iso, oracle):
    """one handle, many extracts (buffers grow and shrink); identical bytes every time"""
    mc = iso.MarchingCubes(64)
    ref = None
    for name in ("torus", "csgA", "sphere03", "torus", "csgA", "torus"):
        mc.extract_device(iso.Sampler(iso_source(name)))
        xyz, idx = mc.copy_out()
        oxyz, oidx, _ = oracle.extract_sdf(64, oracle_prog(name))
        assert mesh_diff(xyz, idx, oxyz, oidx) == "", name
    t = synth(iso, 1, 64, 99)
    for _ in range(20):
        mc.extract_device(iso.DenseGrid(t))
        cur = tuple(a.tobytes() for a in mc.copy_out())
        ref = ref or cur
        assert cur == ref
    mc.close()


def test_extractor_api_paths(iso, oracle):
    """the crate-shaped call: MarchingCubes(size).extract(Sampler(source), IndexedVertices(v, i))"""
    vertices, indices = [], []
    mc = iso.MarchingCubes(32)
    mc.extract(iso.Sampler(iso_source("sphere03")), iso.IndexedVertices(vertices, indices))
    oxyz, oidx, _ = oracle.extract_sdf(32, oracle_prog("sphere03"))
    assert np.array_equal(np.asarray(vertices, np.float32), oxyz) and np.array_equal(np.asarray(indices, np.uint32), oidx)

    class Rec(iso.Extractor):
        def __init__(self):
            self.v, self.i, self.late_vertex = [], [], False

        def extract_vertex(self, v):
            self.late_vertex |= bool(self.i)
            self.v.extend(v)

        def extract_index(self, i):
            self.i.append(i)
    r = Rec()
    mc.extract(iso.Sampler(iso_source("sphere03")), r)
    assert not r.late_vertex and np.array_equal(np.asarray(r.v, np.float32), oxyz) and r.i == oidx.tolist()
    mc.close()


def test_unsupported_and_error_paths(iso, isolib):
    from isosurface_b200 import _lib
    mc = iso.MarchingCubes(16)
    with pytest.raises(_lib.IsomcError) as ei:
        mc.counts()
    assert ei.value.code == _lib.ERR_NO_RESULT
    bad = np.zeros(1, dtype=_lib.NODE_DTYPE)
    bad[0] = (99, 0, 0, 0)
    assert isolib.isomc_extract_sdf(mc._h, bad.ctypes.data, 1) == _lib.ERR_UNSUPPORTED_SOURCE
    assert b"closures" in isolib.isomc_last_error(mc._h)
    two = np.zeros(2, dtype=_lib.NODE_DTYPE)
    two[0] = (1, 1, 0, 0); two[1] = (1, 1, 0, 0)
    assert isolib.isomc_extract_sdf(mc._h, two.ctypes.data, 2) == _lib.ERR_BAD_ARG
    with pytest.raises(ValueError):
        mc.extract_device(iso.DenseGrid(np.zeros((9, 8, 8), np.float32)))
    mc.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_slab_decomposition_on_one_gpu(iso, oracle, world):
    """multi-GPU path simulated serially on one device: concatenated slabs == unsharded == oracle"""
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    size = 72
    t = synth(iso, 1, size, 21)
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, _ = oracle.extract_grid(size, host)
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    ptrs = []
    for r in range(world):
        zf, nl = slab_sample_layers(size, r, world)
        ptrs.append(t.data_ptr() + 4 * zf * size * size)
    totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
    parts = []
    for r, s in enumerate(slabs):
        s.extract(ptrs[r], gathered=totals)
        parts.append(s.copy_out())
    xyz = np.concatenate([p[0] for p in parts])
    idx = np.concatenate([p[1] for p in parts])
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    assert int(totals[:, 0].sum()) == len(oxyz) // 3 and int(totals[:, 2].sum()) == len(oidx) // 3
    for s in slabs:
        s.close()


def test_slab_emit_gathered_on_stream(iso, oracle, isolib):
    """the on-stream variant (device-side base computation from the all-gathered totals)"""
    import torch
    from isosurface_b200 import _lib
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    size, world = 48, 4
    t = synth(iso, 2, size, 0)
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, _ = oracle.extract_grid(size, host)
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
    totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.int64)
    gathered = torch.from_numpy(totals.ravel().copy()).cuda()
    parts = []
    for r, s in enumerate(slabs):
        _lib.check(isolib.isomc_slab_count_grid_device(s._h, C.c_void_p(ptrs[r])), s._h)
        _lib.check(isolib.isomc_slab_emit_gathered(s._h, C.c_void_p(gathered.data_ptr()), r, world), s._h)
        parts.append(s.copy_out())
    assert mesh_diff(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), oxyz, oidx) == ""


def test_layer_counts_and_slabs_of_equal_work(iso, oracle, isolib):
    """isomc_layer_counts reports the per-layer totals of the last extract (whole lattice == the slabs' layers put together), and
    slabs of UNEQUAL thickness cut by sharded.balanced_slabs still concatenate to the reference mesh"""
    from isosurface_b200 import _lib
    from isosurface_b200.sharded import SlabMarchingCubes, balanced_slabs, bases_from_totals, slab_range
    size, world = 48, 4
    t = synth(iso, 3, size, 11)  # sphere union: the surface is unevenly spread over z
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, oact = oracle.extract_grid(size, host)
    mc = iso.MarchingCubes(size)
    mc.extract_device(iso.DenseGrid(t))
    whole = np.zeros(3 * size, np.uint64)
    _lib.check(isolib.isomc_layer_counts(mc._h, whole.ctypes.data), mc._h)
    whole = whole.reshape(size, 3)
    assert [int(x) for x in whole.sum(axis=0)] == [len(oxyz) // 3, len(oidx) // 3, oact]
    ci = oracle.cube_indices(size, host)
    act_ref = ((ci != 0) & (ci != 255)).reshape(size, -1).sum(axis=1)
    assert np.array_equal(whole[:, 2].astype(np.int64), act_ref)
    mc.close()
    slabs = balanced_slabs(size, whole[:, 2], world, active_cell_cost=400.0)  # (a small lattice: exaggerate the surface cost)
    assert slabs != [slab_range(size, r, world) for r in range(world)]
    parts, totals, hs = [], [], []
    for r in range(world):
        s = SlabMarchingCubes(size, r, world, z_range=slabs[r])
        first = slabs[r][0] - (1 if slabs[r][0] > 0 else 0)
        totals.append(s.count(t.data_ptr() + 4 * first * size * size))
        hs.append(s)
    for r, s in enumerate(hs):
        vbase, bbase, _ = bases_from_totals(np.array(totals), r)
        s.emit(vbase, bbase)
        assert np.array_equal(s.layer_active_cells().astype(np.int64), act_ref[slabs[r][0]:slabs[r][1]])
        parts.append(s.copy_out())
        s.close()
    assert mesh_diff(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), oxyz, oidx) == ""


def test_slab_exchange_over_peer_memory_on_one_device(iso, oracle, isolib, monkeypatch):
    """isomc_slab_connect + isomc_slab_emit_exchanged: the totals travel as stores into the ranks' mailboxes, the id offset is
    derived by the waiting kernel.  All ranks on one device here (the mailboxes are plain device pointers); several steps, so that
    both mailbox parities and the step counter are exercised; then a rank that never publishes must time out, not hang."""
    import torch
    from isosurface_b200 import _lib
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    monkeypatch.setenv("ISOMC_EXCHANGE_TIMEOUT_MS", "1500")
    size, world = 48, 4
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    boxes = (C.c_void_p * world)()
    for r, s in enumerate(slabs):
        b = C.c_void_p()
        _lib.check(isolib.isomc_slab_mailbox(s._h, C.byref(b)), s._h)
        boxes[r] = b.value
    for r, s in enumerate(slabs):
        _lib.check(isolib.isomc_slab_connect(s._h, r, world, boxes), s._h)
    for step, seed in enumerate((0, 5, 9)):
        t = synth(iso, 2 if step == 0 else 1, size, seed)
        host = t.cpu().numpy().reshape(size + 1, size, size)
        oxyz, oidx, _ = oracle.extract_grid(size, host)
        ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
        for rep in range(3 if step == 0 else 2):  # (first extract of a handle: no output buffers yet, finish() runs the emission)
            for r, s in enumerate(slabs):
                if rep == 0:  # the two-call form ...
                    _lib.check(isolib.isomc_slab_count_grid_device(s._h, C.c_void_p(ptrs[r])), s._h)
                    _lib.check(isolib.isomc_slab_enqueue_emit_exchanged(s._h), s._h)
                else:         # ... and the one-call form (one launch sequence per rank, replayed from a graph the second time)
                    _lib.check(isolib.isomc_slab_enqueue_extract_grid_exchanged(s._h, C.c_void_p(ptrs[r])), s._h)
            for s in slabs:
                _lib.check(isolib.isomc_finish(s._h), s._h)
        parts = [s.copy_out() for s in slabs]
        assert mesh_diff(np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), oxyz, oidx) == "", step
    # ranks 1.. stay silent: rank 0 gives up after the timeout with an error naming a silent rank
    _lib.check(isolib.isomc_slab_count_grid_device(slabs[0]._h, C.c_void_p(ptrs[0])), slabs[0]._h)
    rc = isolib.isomc_slab_emit_exchanged(slabs[0]._h)
    assert rc == _lib.ERR_NCCL and b"timed out" in isolib.isomc_last_error(slabs[0]._h)
    for s in slabs:
        s.close()


def test_full_size_properties_512(iso, oracle):
    """BASELINE C3 at full size: invariants that do not need the oracle, plus an oracle check on a z-window"""
    size = 512
    t = synth(iso, 1, size, 0x1505F00D)
    mc = iso.MarchingCubes(size)
    nv, nt, na = mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    assert nv == len(xyz) // 3 and nt == len(idx) // 3 and nv > 1_000_000
    assert int(idx.max()) == nv - 1
    ref = np.zeros(nv, bool)
    ref[idx] = True
    assert ref.all()                                   # every vertex referenced
    assert np.isfinite(xyz).all() and xyz.min() >= 0.0 and xyz.max() <= 512 / 511 + 1e-6
    first_use = np.full(nv, np.iinfo(np.int64).max)
    np.minimum.at(first_use, idx, np.arange(len(idx)))
    assert np.all(np.diff(first_use) > 0)              # vertex k is first referenced before vertex k+1 (mesh.rs:240-251)
    # oracle on the first 24 cell layers: exact prefix of the device mesh
    host = t[: 25 * size * size].cpu().numpy().reshape(25, size, size)
    wx, wi, _ = oracle.extract_grid(size, host, z_cells=24)
    assert np.array_equal(idx[:len(wi)], wi) and np.array_equal(xyz[:len(wx)].view(np.uint32), wx.view(np.uint32))
    mc.close()


def test_full_size_slabs_equal_unsharded_1024(iso):
    """BASELINE C4 (1024^3 gyroid): 8 slabs run serially == one unsharded extract, byte for byte"""
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    size, world = 1024, 8
    t = synth(iso, 2, size, 0)
    mc = iso.MarchingCubes(size)
    mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    mc.close()
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
    totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
    vo = io = 0
    for r, s in enumerate(slabs):
        s.extract(ptrs[r], gathered=totals)
        px, pi = s.copy_out()
        assert np.array_equal(xyz[vo:vo + len(px)].view(np.uint32), px.view(np.uint32))
        assert np.array_equal(idx[io:io + len(pi)], pi)
        vo += len(px); io += len(pi)
        s.close()
    assert vo == len(xyz) and io == len(idx)
    facts = mesh_invariants(xyz, idx, closed=False)
    assert facts["directed_edge_dups"] == 0


@pytest.mark.parametrize("env", [{"ISOMC_PATH": "tile"}, {"ISOMC_PATH": "tile", "ISOMC_FILL": "plain"}, {"ISOMC_PATH": "tile", "ISOMC_RING": "2"}])
def test_tile_path_is_bit_identical(iso, oracle, monkeypatch, env):
    """ISOMC_PATH=tile (TMA-staged one-pass count + plane emission; opt-in, profiles/r02_tile_path.md), with the TMA bulk
    copies, with all-thread loads (ISOMC_FILL=plain) and with the two-slot ring: same bytes as the default path and the oracle
    -- grids, slabs, implicit and Directed sources, the streamed host extract"""
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    size = 160
    t = synth(iso, 1, size, 5)
    mc = iso.MarchingCubes(size)
    mc.extract_device(iso.DenseGrid(t))
    ref = [a.tobytes() for a in mc.copy_out()]
    mc.close()
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, _ = oracle.extract_grid(size, host)
    assert ref[1] == oidx.tobytes() and ref[0] == oxyz.tobytes()
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    mp = iso.MarchingCubes(size)          # the environment is read at create
    for _ in range(2):
        mp.extract_device(iso.DenseGrid(t))
        assert [a.tobytes() for a in mp.copy_out()] == ref
    xyz, idx = mp.extract_host(iso.DenseGrid(host))
    assert xyz.tobytes() == ref[0] and idx.tobytes() == ref[1]
    xyz, idx = mp.extract_host(iso.DenseGrid(host))    # second call: the streamed z-chunk pipeline
    assert xyz.tobytes() == ref[0] and idx.tobytes() == ref[1]
    mp.close()
    for size2, kind, seed in ((67, 3, 9), (520, 1, 3)):   # N % 4 != 0; rows of two tiles (halo columns)
        nl = size2 + 1 if size2 < 200 else 13
        t2 = synth(iso, kind, size2, seed, 0, nl)
        h2 = t2.cpu().numpy().reshape(nl, size2, size2)
        wx, wi, _ = oracle.extract_grid(size2, h2, nl - 1)
        s0 = SlabMarchingCubes(size2, 0, 1) if nl == size2 + 1 else None
        if s0 is not None:
            s0.extract(t2.data_ptr())
            px, pi = s0.copy_out()
            s0.close()
            assert mesh_diff(px, pi, wx, wi, POS_TOL) == ""
    world = 3
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
    totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
    parts = []
    for r, s in enumerate(slabs):
        s.extract(ptrs[r], gathered=totals)
        parts.append(s.copy_out())
        s.close()
    assert np.concatenate([p[0] for p in parts]).tobytes() == ref[0] and np.concatenate([p[1] for p in parts]).tobytes() == ref[1]
    for name, n in (("csgA", 64), ("torus_origin", 96)):
        m2 = iso.MarchingCubes(n)
        m2.extract_device(iso.Sampler(iso_source(name)))
        xyz, idx = m2.copy_out()
        m2.close()
        oxyz2, oidx2, _ = oracle.extract_sdf(n, oracle_prog(name))
        assert mesh_diff(xyz, idx, oxyz2, oidx2, POS_TOL) == ""
    md = iso.MarchingCubes(48, distance="directed")
    md.extract_device(iso.Sampler(iso_source("csgB")))
    xyz, idx = md.copy_out()
    md.close()
    oxyz3, oidx3, _ = oracle.extract_sdf_directed(48, oracle_prog("csgB"))
    assert mesh_diff(xyz, idx, oxyz3, oidx3, POS_TOL) == ""


def test_non_multiple_of_four_size_uses_generic_sign_kernel(iso, oracle):
    """N % 4 != 0 (or a misaligned pointer) takes the scalar k_sign path instead of the float4 one"""
    import torch
    for size in (66, 67, 131):
        t = synth(iso, 3, size, 9)
        host = t.cpu().numpy().reshape(size + 1, size, size)
        oxyz, oidx, _ = oracle.extract_grid(size, host)
        mc = iso.MarchingCubes(size)
        mc.extract_device(iso.DenseGrid(t))
        xyz, idx = mc.copy_out()
        assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
        mc.close()
    # misaligned device pointer with N % 4 == 0
    size = 64
    buf = torch.empty(size * size * (size + 1) + 1, dtype=torch.float32, device="cuda:0")
    t = synth(iso, 1, size, 2)
    buf[1:].copy_(t)
    mc = iso.MarchingCubes(size)
    mc.extract_device(iso.DenseGrid(buf.data_ptr() + 4, size=size, on_device=True))
    xyz, idx = mc.copy_out()
    oxyz, oidx, _ = oracle.extract_grid(size, t.cpu().numpy().reshape(size + 1, size, size))
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    mc.close()


def test_randomised_stress_subset():
    """40 cases of tools/gpu_stress.py (random sizes, densities, boundary-aligned fields, random slab counts);
    the full 400-case run is recorded in profiles/r01_sanitizer.txt"""
    import subprocess
    import sys
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, str(root / "tools" / "gpu_stress.py"), "40", "123"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "40 cases, 0 failures" in r.stdout


def test_streamed_host_to_host_extract(iso, oracle):
    """isomc_extract_grid_host_to: z-chunked copy-in / kernels / copy-out pipeline gives the oracle's bytes, on the
    plain first call, on the streamed later calls, when the mesh changes size, and reports too-small caller buffers"""
    import torch
    from isosurface_b200 import _lib
    size = 96
    fields = [synth(iso, 1, size, s).cpu().numpy().reshape(size + 1, size, size) for s in (3, 4)]
    fields.append(synth(iso, 3, size, 5).cpu().numpy().reshape(size + 1, size, size))      # much smaller mesh
    rng = np.random.default_rng(0)
    fields.append(rng.standard_normal((size + 1, size, size)).astype(np.float32))         # much larger mesh: regrow + re-run
    want = [oracle.extract_grid(size, f) for f in fields]
    mc = iso.MarchingCubes(size)
    for rep in range(2):
        for f, (oxyz, oidx, _) in zip(fields, want):
            xyz, idx = mc.extract_host(iso.DenseGrid(f))
            assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    # caller-owned pinned buffers, exactly sized, streamed path
    f, (oxyz, oidx, _) = fields[0], want[0]
    hg = torch.from_numpy(f).pin_memory()
    hx = torch.empty(len(oxyz), dtype=torch.float32).pin_memory()
    hi = torch.empty(len(oidx), dtype=torch.int32).pin_memory()
    lib = _lib.load()
    for _ in range(2):
        hx.fill_(-1); hi.fill_(-1)
        _lib.check(lib.isomc_extract_grid_host_to(mc._h, C.c_void_p(hg.data_ptr()), C.c_void_p(hx.data_ptr()), len(oxyz) // 3,
                                                  C.c_void_p(hi.data_ptr()), len(oidx) // 3), mc._h)
        assert mesh_diff(hx.numpy(), hi.numpy().view(np.uint32), oxyz, oidx, POS_TOL) == ""
    # one triangle short: reported, result stays fetchable
    rc = lib.isomc_extract_grid_host_to(mc._h, C.c_void_p(hg.data_ptr()), C.c_void_p(hx.data_ptr()), len(oxyz) // 3,
                                        C.c_void_p(hi.data_ptr()), len(oidx) // 3 - 1)
    assert rc == _lib.ERR_BUFFER_TOO_SMALL
    xyz, idx = mc.copy_out()
    assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    # Extractor API with a host grid goes through the same call
    sink = iso.ArrayMesh()
    mc.extract(iso.DenseGrid(fields[1]), sink)
    assert mesh_diff(sink.vertices.ravel(), sink.indices.ravel(), want[1][0], want[1][1], POS_TOL) == ""
    mc.close()


def test_point_cloud_matches_oracle(iso, oracle):
    """PointCloud(size).extract: same points, same order, same bits as the restated reference (point_cloud.rs:50-63)"""
    for name, size in (("sphere03", 32), ("csgA", 64), ("torus_origin", 128)):
        want = oracle.point_cloud_sdf(size, oracle_prog(name))
        pc = iso.PointCloud(size)
        verts = []
        pc.extract(iso.Sampler(iso_source(name)), iso.OnlyVertices(verts))
        assert np.asarray(verts, np.float32).tobytes() == want.tobytes()
        nv, nt, na = pc.counts()
        assert (nv, nt, na) == (len(want) // 3, 0, len(want) // 3)
        pc.close()
    rng = np.random.default_rng(3)
    for size in (2, 3, 33, 70, 130):
        f = rng.standard_normal((size + 1, size, size)).astype(np.float32)
        want = oracle.point_cloud_grid(size, f)
        pc = iso.PointCloud(size)
        for _ in range(2):
            sink = iso.ArrayMesh()
            pc.extract(iso.DenseGrid(f), sink)
            assert sink.vertices.tobytes() == want.tobytes() and sink.indices.size == 0
        pc.close()
    # device-resident lattice, a mesh extract on the same handle type afterwards is unaffected
    size = 160
    t = synth(iso, 2, size, 1)
    want = oracle.point_cloud_grid(size, t.cpu().numpy().reshape(size + 1, size, size))
    pc = iso.PointCloud(size)
    pc.extract_device(iso.DenseGrid(t))
    assert pc.copy_out()[0].tobytes() == want.tobytes()
    pc.close()
    assert iso.PointCloud(1).extract_device(iso.Sampler(iso_source("sphere03")))[0] == 0


def test_directed_point_cloud_matches_oracle(iso, oracle):
    """PointCloud<Directed> over implicit trees: bit-identical to the restatement; a dense lattice is refused"""
    for name, size in (("sphere03", 24), ("csgA", 40), ("csgB", 33), ("torus_origin", 32)):
        want = oracle.point_cloud_sdf_directed(size, oracle_prog(name))
        pc = iso.PointCloud(size, distance="directed")
        pts = []
        pc.extract(iso.Sampler(iso_source(name)), iso.OnlyVertices(pts))
        assert np.array_equal(np.asarray(pts, np.float32).view(np.uint32), want.view(np.uint32)), (name, size)
        with pytest.raises(TypeError):
            pc.extract_device(iso.DenseGrid(np.zeros((size + 1, size, size), np.float32)))
        pc.close()


def test_interleaved_normals_match_oracle(iso, oracle):
    """IndexedInterleavedNormals + CentralDifference (examples/sampler.rs:79-103): positions and central-difference
    normals, evaluated on the device, bit-identical to the restated reference"""
    O = oracle
    cases = {
        "csgA": (iso.Union(iso.Difference(iso.Sphere(.25), iso.RectangularPrism((.2, .2, .2))), iso.Cylinder(.02, .25)),
                 [(O.SPHERE, .25), (O.PRISM, .2, .2, .2), (O.DIFFERENCE,), (O.CYLINDER, .02, .25), (O.UNION,)]),
        "csgB": (iso.Intersection(iso.Sphere(.3), iso.RectangularPrism((.2, .2, .2))),
                 [(O.SPHERE, .3), (O.PRISM, .2, .2, .2), (O.INTERSECTION,)]),
        "torus": (iso.Torus(.25, .1), [(O.TORUS, .25, .1)]),
    }
    for name, (tree, nodes) in cases.items():
        for eps in (0.000001, 0.001):
            size = 64
            src = iso.Translate(.5, iso.CentralDifference(tree, eps))          # DemoSource(CentralDifference(tree))
            sampler = iso.Sampler(src)
            verts, inds = [], []
            mc = iso.MarchingCubes(size)
            mc.extract(sampler, iso.IndexedInterleavedNormals(verts, inds, sampler))
            mc.close()
            oxyz, oidx, _ = O.extract_sdf(size, oracle_prog(name))
            want = O.interleaved_normals_cd(O.program(nodes), oxyz, eps, [(.5, .5, .5)])
            got = np.asarray(verts, np.float32).reshape(-1, 6)
            assert np.asarray(inds, np.uint32).tobytes() == oidx.tobytes()
            assert got[:, :3].tobytes() == oxyz.tobytes()
            same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
            assert same.all(), "%s eps=%g: %d normal components differ" % (name, eps, int((~same).sum()))
    with pytest.raises(TypeError):
        iso.IndexedInterleavedNormals([], [], iso.Sampler(iso.Sphere(.3)))     # no CentralDifference: not a device path


def test_directed_marching_cubes_matches_oracle(iso, oracle):
    """MarchingCubes<Directed> (reference src/distance.rs:72-104 + the shapes' sample_vector): indices and vertex bits
    identical to the restated reference; the restatement's sample_vector is pinned by the reference's own unit tests"""
    for name, size in (("sphere03", 32), ("torus", 64), ("csgA", 64), ("csgB", 64), ("nested", 100), ("cylinder", 65),
                       ("prism", 33), ("torus_origin", 128), ("csgA", 256)):
        oxyz, oidx, oact = oracle.extract_sdf_directed(size, oracle_prog(name))
        mc = iso.MarchingCubes(size, distance="directed")
        for _ in range(2):
            sink = iso.ArrayMesh()
            mc.extract(iso.Sampler(iso_source(name)), sink)
            assert mesh_diff(sink.vertices.ravel(), sink.indices.ravel(), oxyz, oidx, POS_TOL) == "", (name, size)
            assert mc.counts()[2] == oact
        mc.close()
    with pytest.raises(TypeError):
        mc = iso.MarchingCubes(8, distance="directed")
        mc.extract_device(iso.DenseGrid(np.zeros((9, 8, 8), np.float32)))


def test_chunked_driver_overlaps_small_extracts(iso, oracle):
    """ChunkedMarchingCubes: many 32^3 chunks in flight on independent handles/streams; every chunk's mesh is the plain
    API's (= the oracle's), in submission order"""
    size = 32
    offsets = [(0.3 + 0.05 * i, 0.5, 0.45 + 0.01 * i) for i in range(13)]
    sources = [iso.Translate(o, iso.Union(iso.Sphere(0.2), iso.Torus(0.25, 0.08))) for o in offsets]
    want = []
    for o in offsets:
        prog = oracle.program([(oracle.TRANSLATE_PUSH,) + o, (oracle.SPHERE, .2), (oracle.TORUS, .25, .08), (oracle.UNION,), (oracle.TRANSLATE_POP,)])
        want.append(oracle.extract_sdf(size, prog))
    drv = iso.ChunkedMarchingCubes(size, n_inflight=4)
    for _ in range(2):
        got = drv.extract_many([iso.Sampler(s) for s in sources])
        assert len(got) == len(want)
        for (xyz, idx), (oxyz, oidx, _) in zip(got, want):
            assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    seen = []
    drv.extract_many(sources[:5], deliver=lambda i, xyz, idx: seen.append((i, len(idx))))
    assert [i for i, _ in seen] == [0, 1, 2, 3, 4] and all(n == len(w[1]) for (_, n), w in zip(seen, want))
    drv.close()


def test_batched_chunks_match_single_extracts(iso, oracle):
    """BatchedMarchingCubes / isomc_extract_sdf_batch: B chunks through one kernel sequence; each chunk's mesh is byte
    for byte the single-chunk oracle mesh (chunk-local indices), incl. empty chunks, partial batches and handle reuse"""
    for size, cap in ((32, 16), (9, 5), (48, 7)):
        offsets = [(0.3 + 0.05 * i, 0.5, 0.45 + 0.01 * i) for i in range(13)] + [(5.0, 5.0, 5.0)]  # the last chunk is empty
        sources = [iso.Translate(o, iso.Union(iso.Sphere(0.2), iso.Torus(0.25, 0.08))) for o in offsets]
        sources.append(iso.Difference(iso.Sphere(0.3), iso.Translate((0.5, 0.5, 0.5), iso.Sphere(0.4))))
        want = []
        for o in offsets:
            prog = oracle.program([(oracle.TRANSLATE_PUSH,) + o, (oracle.SPHERE, .2), (oracle.TORUS, .25, .08), (oracle.UNION,),
                                   (oracle.TRANSLATE_POP,)])
            want.append(oracle.extract_sdf(size, prog))
        from isosurface_b200.source import encode_program
        want.append(oracle.extract_sdf(size, encode_program(sources[-1])))
        drv = iso.BatchedMarchingCubes(size, n_chunks=cap)
        for _ in range(2):
            got = drv.extract_many([iso.Sampler(s) for s in sources])
            assert len(got) == len(want)
            for b, ((xyz, idx), (oxyz, oidx, _)) in enumerate(zip(got, want)):
                assert mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == "", (size, cap, b)
        # the plain single-chunk handle gives the same bytes
        mc = iso.MarchingCubes(size)
        mc.extract_device(iso.Sampler(sources[3]))
        xyz, idx = mc.copy_out()
        assert mesh_diff(xyz, idx, got[3][0], got[3][1], POS_TOL) == ""
        mc.close()
        drv.close()
    from isosurface_b200 import _lib
    with pytest.raises(_lib.IsomcError):
        iso.BatchedMarchingCubes(1, n_chunks=4)


def test_batched_directed_chunks_match_oracle(iso, oracle):
    """isomc_extract_sdf_batch_directed: MarchingCubes<Directed> per chunk through one kernel sequence; every chunk = the oracle's
    Directed mesh of that tree (incl. a partly filled batch: the padding lattices must stay empty for vector distances too)"""
    for size, cap in ((24, 5), (33, 3)):
        names = ["sphere03", "torus", "csgA", "csgB", "sphere05_origin"][:cap]
        want = [oracle.extract_sdf_directed(size, oracle_prog(n)) for n in names]
        drv = iso.BatchedMarchingCubes(size, n_chunks=cap + 2, distance="directed")
        for _ in range(2):
            xyz, idx, vo, to = drv.extract_batch([iso.Sampler(iso_source(n)) for n in names])
            assert len(vo) == cap + 1
            for b in range(cap):
                oxyz, oidx, _ = want[b]
                cx, ci = xyz[3 * int(vo[b]):3 * int(vo[b + 1])], idx[3 * int(to[b]):3 * int(to[b + 1])]
                assert mesh_diff(cx, ci, oxyz, oidx, POS_TOL) == "", (size, names[b])
            assert len(xyz) == 3 * int(vo[cap]) and len(idx) == 3 * int(to[cap])  # nothing in the padding lattices
        with pytest.raises(TypeError):
            drv.extract_grids(np.zeros((1, size + 1, size, size), np.float32))
        drv.close()


def test_batched_dense_chunks_match_single_extracts(iso, oracle):
    """isomc_extract_grid_batch_{host,device}: dense size^3 chunks (a voxel world cut into chunks) through one kernel sequence; every
    chunk's mesh is the oracle's mesh of that lattice alone, incl. empty chunks, a partly filled host batch and handle reuse"""
    import torch
    rng = np.random.default_rng(11)
    for size, cap in ((16, 6), (33, 4), (40, 3)):
        shapes = ["sphere03", "torus", "csgA", "sphere05_origin"]
        grids = [oracle.fill_grid_sdf(size, oracle_prog(shapes[b % 4])) if b % 2 == 0 else rng.standard_normal((size + 1, size, size)).astype(np.float32)
                 for b in range(cap)]
        grids[-1] = np.full((size + 1, size, size), 2.0, np.float32)  # an empty chunk
        want = [oracle.extract_grid(size, g) for g in grids]
        drv = iso.BatchedMarchingCubes(size, n_chunks=cap)
        stacked = np.stack([g.reshape(size + 1, size, size) for g in grids])
        dev = torch.from_numpy(stacked).cuda()
        for src, n in ((stacked, cap), (dev, cap), (stacked[:cap - 1], cap - 1), (dev, cap)):
            xyz, idx, vo, to = drv.extract_grids(src)
            assert len(vo) == n + 1
            for b in range(n):
                oxyz, oidx, _ = want[b]
                cx, ci = xyz[3 * int(vo[b]):3 * int(vo[b + 1])], idx[3 * int(to[b]):3 * int(to[b + 1])]
                assert mesh_diff(cx, ci, oxyz, oidx, POS_TOL) == "", (size, cap, b, n)
        drv.close()


# ---- full-size parity (SURVEY 8c / VERDICT r01): committed oracle hashes of the benchmark fields, made on a B200 box by
# tools/gen_golden_full.py (device-generated field bytes -> CPU oracle in lean mode); plus one live full-size oracle run
FULL_PATH = Path(__file__).parent / "golden" / "full_hashes.json"
FULL = json.loads(FULL_PATH.read_text()) if FULL_PATH.exists() else {}


def _golden_field(iso, name):
    if name not in FULL:
        pytest.skip("tests/golden/full_hashes.json has no entry for %s (run tools/gen_golden_full.py on the GPU box)" % name)
    g = FULL[name]
    t = synth(iso, g["kind"], g["size"], g["seed"], 0, g["z_cells"] + 1)
    return g, t


@pytest.mark.parametrize("name", ["fbm512", "gyroid1024"])
def test_full_size_mesh_matches_committed_oracle_hashes(iso, name):
    """BASELINE C3 / C4 at full size: the WHOLE device mesh (every layer, the last ones included) hashes to what the CPU
    oracle produced from the same field bytes"""
    g, t = _golden_field(iso, name)
    mc = iso.MarchingCubes(g["size"])
    nv, nt, na = mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    mc.close()
    assert (na, nv, nt) == (g["active_cells"], g["vertices"], g["triangles"])
    assert sha(idx, "<u4") == g["sha_i"], "index stream differs from the oracle's"
    assert sha(xyz, "<f4") == g["sha_v"], "vertex stream differs from the oracle's"


@pytest.mark.parametrize("name", ["fbm644_z32", "fbm812_z32", "fbm1024_z24", "spheres2048_z64"])
def test_window_matches_committed_oracle_hashes(iso, isolib, name):
    """the workloads the scaling runs time (fbm644 / fbm812 / fbm1024) and C5: the first cell layers against the oracle"""
    from isosurface_b200 import _lib
    g, t = _golden_field(iso, name)
    h = C.c_void_p()
    _lib.check(isolib.isomc_slab_create(g["size"], 0, g["z_cells"], 0, C.byref(h)))
    _lib.check(isolib.isomc_slab_count_grid_device(h, C.c_void_p(t.data_ptr())), h)
    _lib.check(isolib.isomc_slab_emit(h, 0, 0), h)
    v, tr, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _lib.check(isolib.isomc_counts(h, C.byref(v), C.byref(tr), C.byref(a)), h)
    xyz, idx = np.empty(3 * v.value, np.float32), np.empty(3 * tr.value, np.uint32)
    _lib.check(isolib.isomc_copy_out(h, xyz.ctypes.data, idx.ctypes.data), h)
    isolib.isomc_destroy(h)
    assert (a.value, v.value, tr.value) == (g["active_cells"], g["vertices"], g["triangles"])
    assert sha(idx, "<u4") == g["sha_i"] and sha(xyz, "<f4") == g["sha_v"]


@pytest.mark.parametrize("size,z_cells,kind,seed", [(2112, 3, 1, 5), (2600, 2, 3, 9)])
def test_rows_wider_than_a_warp_pass(iso, isolib, oracle, size, z_cells, kind, seed):
    """lattices wider than 2049 samples: a cell row has more than 64 segments, so the counting kernel runs its WIDE form (a pass is a
    64-segment chunk of one row, rows closed across passes).  A few cell layers of such a lattice against the live oracle."""
    from isosurface_b200 import _lib
    t = synth(iso, kind, size, seed, 0, z_cells + 1)
    host = t.cpu().numpy().reshape(z_cells + 1, size, size)
    oxyz, oidx, oact = oracle.extract_grid(size, host, z_cells=z_cells)
    assert oact > 1000
    h = C.c_void_p()
    _lib.check(isolib.isomc_slab_create(size, 0, z_cells, 0, C.byref(h)))
    for _ in range(2):  # (second extract: inline emission, graph replay)
        _lib.check(isolib.isomc_slab_count_grid_device(h, C.c_void_p(t.data_ptr())), h)
        _lib.check(isolib.isomc_slab_emit(h, 0, 0), h)
        v, tr, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(isolib.isomc_counts(h, C.byref(v), C.byref(tr), C.byref(a)), h)
        xyz, idx = np.empty(3 * v.value, np.float32), np.empty(3 * tr.value, np.uint32)
        _lib.check(isolib.isomc_copy_out(h, xyz.ctypes.data, idx.ctypes.data), h)
        assert a.value == oact and mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    isolib.isomc_destroy(h)


def test_fbm512_whole_mesh_against_live_oracle(iso, oracle):
    """BASELINE C3: the whole 512^3 fBm mesh (6.9 M vertices, 13.8 M triangles) against a LIVE lean oracle run on the same
    field bytes -- memcmp on both streams (tens of seconds of CPU time)"""
    size = 512
    t = synth(iso, 1, size, 0x1505F00D)
    mc = iso.MarchingCubes(size)
    mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    mc.close()
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, _ = oracle.extract_grid(size, host, size, oracle.LEAN)
    assert len(idx) == len(oidx) and np.array_equal(idx, oidx)
    assert len(xyz) == len(oxyz) and np.array_equal(xyz.view(np.uint32), oxyz.view(np.uint32))


def test_full_size_2048_slabs_equal_unsharded(iso):
    """BASELINE C5 (2048^3 sphere union, 34 GB): the unsharded extract -- sample offsets beyond 2^32, 8.6 G cells in one
    handle -- equals the 8-slab decomposition byte for byte, and the size-independent invariants hold"""
    import torch
    if torch.cuda.get_device_properties(0).total_memory < 60 * (1 << 30):
        pytest.skip("needs ~45 GB of device memory")
    from isosurface_b200.sharded import SlabMarchingCubes, slab_sample_layers
    size, world = 2048, 8
    t = synth(iso, 3, size, 0x5EEDBA11)
    mc = iso.MarchingCubes(size)
    nv, nt, na = mc.extract_device(iso.DenseGrid(t))
    xyz, idx = mc.copy_out()
    mc.close()
    assert nv == len(xyz) // 3 and nt == len(idx) // 3 and nt > 50_000_000
    assert int(idx.max()) == nv - 1
    ref = np.zeros(nv, bool)
    ref[idx] = True
    assert ref.all()                                   # every vertex referenced
    del ref
    assert np.isfinite(xyz).all() and xyz.min() >= 0.0 and xyz.max() <= 2048 / 2047 + 1e-6
    head = idx[: 30_000_000]
    first_use = np.full(int(head.max()) + 1, np.iinfo(np.int64).max)
    np.minimum.at(first_use, head, np.arange(len(head)))
    assert np.all(np.diff(first_use) > 0)              # vertex k is first referenced before vertex k+1 (mesh.rs:240-251)
    del first_use, head
    slabs = [SlabMarchingCubes(size, r, world) for r in range(world)]
    ptrs = [t.data_ptr() + 4 * slab_sample_layers(size, r, world)[0] * size * size for r in range(world)]
    totals = np.array([s.count(p) for s, p in zip(slabs, ptrs)], dtype=np.uint64)
    vo = io = 0
    for r, s in enumerate(slabs):
        s.extract(ptrs[r], gathered=totals)
        px, pi = s.copy_out()
        assert np.array_equal(xyz[vo:vo + len(px)].view(np.uint32), px.view(np.uint32)), "slab %d vertices" % r
        assert np.array_equal(idx[io:io + len(pi)], pi), "slab %d indices" % r
        vo += len(px); io += len(pi)
        s.close()
    assert vo == len(xyz) and io == len(idx)


@pytest.mark.parametrize("world", [1, 3, 4])
def test_sharded_api_on_one_device(iso, oracle, world):
    """isomc_sharded_*: all slabs on device 0 (the exchange is then a stream-ordered device copy): == unsharded == oracle;
    a second extract of a different field re-uses the handles"""
    from isosurface_b200.sharded import ShardedMarchingCubes
    size = 80
    sh = ShardedMarchingCubes(size, [0] * world)
    assert not sh.uses_nccl
    for kind, seed in ((1, 31), (3, 7), (1, 32)):
        t = synth(iso, kind, size, seed)
        host = t.cpu().numpy().reshape(size + 1, size, size)
        oxyz, oidx, oact = oracle.extract_grid(size, host)
        ptrs = [t.data_ptr() + 4 * sh.slab(r)[2] * size * size for r in range(world)]
        nv, nt, na = sh.extract_grid(ptrs)
        xyz, idx = sh.copy_out()
        assert na == oact and mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    nv, nt, na = sh.extract_sdf(iso.Sampler(iso_source("csgA")))
    xyz, idx = sh.copy_out()
    oxyz, oidx, oact = oracle.extract_sdf(size, oracle_prog("csgA"))
    assert na == oact and mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == ""
    sh.close()


def test_directed_marching_cubes_on_slabs(iso, oracle):
    """MarchingCubes<Directed> sharded into z-slabs (isomc_sharded_extract_sdf_directed): the concatenation is the unsharded
    Directed mesh of the restatement"""
    from isosurface_b200.sharded import ShardedMarchingCubes
    for name, size, world in (("csgA", 40, 3), ("torus", 33, 4)):
        oxyz, oidx, oact = oracle.extract_sdf_directed(size, oracle_prog(name))
        sh = ShardedMarchingCubes(size, [0] * world)
        for _ in range(2):
            nv, nt, na = sh.extract_sdf(iso.Sampler(iso_source(name)), distance="directed")
            xyz, idx = sh.copy_out()
            assert na == oact and mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == "", (name, world)
        sh.close()


def test_sharded_api_nccl_all_devices(iso, oracle):
    """isomc_sharded_* over every GPU of the box, each slab's lattice on its own device; both forms of the totals exchange"""
    import torch
    from isosurface_b200 import _lib
    from isosurface_b200.sharded import ShardedMarchingCubes
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs (the single-device mode is covered by test_sharded_api_on_one_device)")
    size = 96
    t = synth(iso, 1, size, 77)
    host = t.cpu().numpy().reshape(size + 1, size, size)
    oxyz, oidx, oact = oracle.extract_grid(size, host)
    for mode in ("peer", "nccl"):  # peer stores into mailboxes over NVLink (default), then the NCCL all-gather
        if mode == "nccl":
            os.environ["ISOMC_EXCHANGE"] = "nccl"
        try:
            sh = ShardedMarchingCubes(size, list(range(n)))
        finally:
            os.environ.pop("ISOMC_EXCHANGE", None)
        assert sh.uses_nccl == (mode == "nccl") and sh.uses_peer_memory == (mode == "peer")
        parts = []
        for r in range(n):
            zb, ze, first, nl = sh.slab(r)
            parts.append(torch.from_numpy(host[first:first + nl].copy()).to("cuda:%d" % r))
        for _ in range(3):
            nv, nt, na = sh.extract_grid([p.data_ptr() for p in parts])
            xyz, idx = sh.copy_out()
            assert na == oact and mesh_diff(xyz, idx, oxyz, oidx, POS_TOL) == "", mode
        sh.close()
