"""Host model of the active-cell-list path (tests/list_model.cu) against the oracle.

The model executes the kernels' own source (isomc_cell.cuh: count_list_warp, emit_cell) on the CPU with
emulated warps, so the list format, the neighbour lookups, the warp-level scans and the block allocator are
checked bit for bit without a GPU.  The GPU parity tests then only have to confirm the launch plumbing.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import mesh_diff, oracle_prog

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "list_model.cu"
SO = ROOT / "tests" / "_build" / "liblist_model.so"
DEPS = [SRC] + [ROOT / "isosurface_b200" / "csrc" / n for n in ("isomc_cell.cuh", "isomc_device.cuh", "isomc_tables.h", "isomc_case_table.h")]


def _load(so, extra=()):
    so.parent.mkdir(exist_ok=True)
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in DEPS):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-arch=sm_100a", "-DISOMC_HOST_MODEL", *extra, "-Xcompiler",
                        "-ffp-contract=off,-fPIC,-fno-fast-math", "-shared", "-o", str(so), str(SRC)], check=True)
    lib = C.CDLL(str(so))
    lib.list_model_extract.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


@pytest.fixture(scope="module")
def model_long_tasks():
    """the same model with the 64-pass tasks of very large lattices switched on for anything above 64 passes"""
    return _load(SO.with_name("liblist_model_long.so"), ("-DISOMC_COUNT_LONG_TASK_AT=64u",))


@pytest.fixture(scope="module")
def model():
    lib = _load(SO)
    lib.list_model_extract.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.list_model_extract_directed.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                                C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.list_model_sample_vector.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


def run_model(lib, size, grid, z_begin=0, z_end=None, n_warps=7, seed=1, vofs=0, cap_blocks=4096, cap_v=None, cap_t=None):
    z_end = size if z_end is None else z_end
    ghost = 1 if z_begin > 0 else 0
    slab = np.ascontiguousarray(grid[z_begin - ghost:z_end + 1], dtype=np.float32)
    cap_v = 4 * slab.size if cap_v is None else cap_v
    cap_t = 4 * slab.size if cap_t is None else cap_t
    xyz = np.full(3 * cap_v, np.nan, np.float32)
    idx = np.full(3 * cap_t, 0xFFFFFFFF, np.uint32)
    tot = np.zeros(6, np.uint64)
    rc = lib.list_model_extract(size, z_begin, z_end, slab.ctypes.data, n_warps, seed, vofs, cap_blocks, xyz.ctypes.data, cap_v,
                                idx.ctypes.data, cap_t, tot.ctypes.data)
    assert rc in (0, 1), "model failed rc=%d" % rc
    return rc, xyz[:3 * int(tot[0])], idx[:3 * int(tot[2])], [int(t) for t in tot]


def noise(size, seed, z_layers=None):
    rng = np.random.default_rng(seed)
    z_layers = size + 1 if z_layers is None else z_layers
    return rng.standard_normal((z_layers, size, size)).astype(np.float32)


@pytest.mark.parametrize("name,size", [("sphere03", 32), ("torus", 40), ("csgA", 48), ("sphere05_origin", 33), ("torus_origin", 64)])
def test_model_matches_oracle_on_shapes(model, oracle, name, size):
    prog = oracle_prog(name)
    grid = oracle.fill_grid_sdf(size, prog)
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    rc, xyz, idx, tot = run_model(model, size, grid)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


@pytest.mark.parametrize("size,seed,n_warps", [(2, 1, 1), (3, 2, 3), (17, 3, 5), (33, 4, 9), (34, 5, 2), (65, 6, 11), (70, 7, 64)])
def test_model_matches_oracle_on_noise(model, oracle, size, seed, n_warps):
    """white noise: ~every cell active, every boundary-ownership case, dense segments (32 cells, 5 triangles each)"""
    grid = noise(size, seed)
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    rc, xyz, idx, tot = run_model(model, size, grid, n_warps=n_warps, seed=seed)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


def test_model_wide_rows(model, oracle):
    """rows of more than 32 segments take the chunked path (N > 1025): a thin z window of a 1060-wide lattice"""
    size, zc = 1060, 2
    rng = np.random.default_rng(11)
    grid = rng.standard_normal((zc + 1, size, size)).astype(np.float32)
    grid[:, :, 200:900] = np.abs(grid[:, :, 200:900])  # long empty stretches: chunks without cells
    grid[:, 300:310, 100:1059] = -np.abs(grid[:, 300:310, 100:1059])
    oxyz, oidx, oact = oracle.extract_grid(size, grid, z_cells=zc)
    rc, xyz, idx, tot = run_model(model, size, grid, z_end=zc, n_warps=13, cap_blocks=40000)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


@pytest.mark.parametrize("size,zc,n_warps", [(1060, 1, 3), (1100, 1, 1), (2080, 1, 5)])
def test_model_wide_rows_dense(model, oracle, size, zc, n_warps):
    """dense wide rows: more than 32 queued segments per row, so rows are flushed piecewise (open-row carries, the
    row-end cases 'last segment still queued' and 'everything already flushed'); every 7th row is empty"""
    rng = np.random.default_rng(size)
    grid = rng.standard_normal((zc + 1, size, size)).astype(np.float32)
    grid[:, ::7, :] = 1.0
    grid[:, 5, : size // 2] = 1.0  # a row whose queued segments start mid-row
    oxyz, oidx, oact = oracle.extract_grid(size, grid, z_cells=zc)
    rc, xyz, idx, tot = run_model(model, size, grid, z_end=zc, n_warps=n_warps, cap_blocks=60000)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


def test_model_sparse_rows_batch_across_passes(model, oracle):
    """a sparse field: few non-uniform segments per pass, so one flush window mixes segments of many rows and passes"""
    size = 96
    prog = oracle_prog("sphere03")
    grid = oracle.fill_grid_sdf(size, prog)
    grid[40:44, 10:12, 3:90] *= -1.0  # some scattered sign flips
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    for n_warps in (1, 4, 29):
        rc, xyz, idx, tot = run_model(model, size, grid, n_warps=n_warps, seed=n_warps)
        assert rc == 0 and tot[3] == oact
        assert mesh_diff(xyz, idx, oxyz, oidx) == ""


def test_model_slabs_concatenate(model, oracle):
    """two slabs with the ghost layer and the all-gathered bases give the unsharded mesh"""
    size, cut = 24, 11
    grid = noise(size, 21)
    oxyz, oidx, _ = oracle.extract_grid(size, grid)
    rc0, x0, i0, t0 = run_model(model, size, grid, 0, cut)
    # rank 1: vertex base = V0; boundary base = vertices rank 0 created before its last cell layer
    rc1, x1, i1, t1 = run_model(model, size, grid, cut, size, vofs=t0[1])
    assert rc0 == 0 and rc1 == 0
    assert mesh_diff(np.concatenate([x0, x1]), np.concatenate([i0, i1]), oxyz, oidx) == ""


@pytest.mark.parametrize("size,seed,n_warps", [(70, 7, 3), (33, 4, 9), (1060, 2, 5)])
def test_model_long_tasks(model_long_tasks, oracle, size, seed, n_warps):
    """the 64-pass tasks (lattices with more than 2^20 passes on the GPU) on small inputs"""
    zc = 1 if size > 1000 else None
    grid = noise(size, seed, z_layers=(zc + 1) if zc else None)
    oxyz, oidx, oact = oracle.extract_grid(size, grid, z_cells=zc)
    rc, xyz, idx, tot = run_model(model_long_tasks, size, grid, z_end=zc, n_warps=n_warps, seed=seed, cap_blocks=40000)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


def test_model_list_overflow_is_reported(model, oracle):
    size = 33
    grid = noise(size, 5)
    _, _, oact = oracle.extract_grid(size, grid)
    rc, _, _, tot = run_model(model, size, grid, cap_blocks=8)
    assert rc == 1 and tot[3] == oact and tot[4] > 8     # totals stay right, the host can size the list and re-run
    rc, xyz, idx, tot2 = run_model(model, size, grid, cap_blocks=tot[4])
    assert rc == 0 and tot2[:4] == tot[:4]


DIRECTED_SHAPES = ["sphere03", "torus", "csgA", "csgB", "prism", "cylinder", "nested", "torus_origin"]


@pytest.mark.parametrize("name", DIRECTED_SHAPES)
def test_device_vector_evaluator_matches_oracle_bitwise(model, oracle, name):
    """sdf_eval_vec (the device's VectorSource::sample_vector, run on the host) against the C restatement: random points
    plus lattice points (exact zeros, points on the axes), every component bit for bit (NaN where both are NaN)"""
    prog = oracle_prog(name)
    rng = np.random.default_rng(5)
    pts = rng.uniform(-0.2, 1.2, (20000, 3)).astype(np.float32)
    lat = np.stack(np.meshgrid(*[np.arange(17, dtype=np.float32) / np.float32(16)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pts = np.concatenate([pts, lat, lat - np.float32(0.5)]).astype(np.float32)
    want = oracle.sample_sdf_vector(prog, pts)
    got = np.zeros_like(want)
    assert model.list_model_sample_vector(prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), got.ctypes.data) == 0
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), "%d components differ, first at point %s" % (int((~same).sum()), pts[np.argwhere(~same)[0][0]])


@pytest.mark.parametrize("name,size", [("sphere03", 32), ("torus", 40), ("csgA", 48), ("csgB", 33), ("torus_origin", 64), ("nested", 36)])
def test_model_directed_extract_matches_oracle(model, oracle, name, size):
    """MarchingCubes<Directed>: the list kernels' source with the Directed source against the restated reference"""
    prog = oracle_prog(name)
    oxyz, oidx, oact = oracle.extract_sdf_directed(size, prog)
    cap_v, cap_t = len(oxyz) // 3 + 64, len(oidx) // 3 + 64
    xyz = np.full(3 * cap_v, np.nan, np.float32)
    idx = np.full(3 * cap_t, 0xFFFFFFFF, np.uint32)
    tot = np.zeros(6, np.uint64)
    rc = model.list_model_extract_directed(size, prog.ctypes.data, len(prog), 5, 3, 4096, xyz.ctypes.data, cap_v, idx.ctypes.data, cap_t,
                                           tot.ctypes.data)
    assert rc == 0 and int(tot[3]) == oact
    assert mesh_diff(xyz[:3 * int(tot[0])], idx[:3 * int(tot[2])], oxyz, oidx) == ""


@pytest.mark.parametrize("size,batch,n_warps", [(9, 3, 3), (33, 4, 5), (40, 2, 7), (66, 2, 4)])
def test_model_batched_chunks_equal_single_extracts(model, oracle, size, batch, n_warps):
    """isomc_extract_sdf_batch layout: `batch` lattices stacked in z, one count / scan / emit, chunk-local ids; every chunk must be
    byte for byte what the oracle returns for that lattice alone (the dead cell layer between two lattices contributes nothing)"""
    model.list_model_extract_batch.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                               C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
    names = ["sphere03", "torus", "csgA", "sphere05_origin"]
    grids = [oracle.fill_grid_sdf(size, oracle_prog(names[b % 4])) if b % 2 == 0 else noise(size, 100 + b) for b in range(batch)]
    grids[-1] = np.full_like(grids[0], 1.0)  # an empty chunk at the end
    stacked = np.ascontiguousarray(np.concatenate([g.reshape(size + 1, size, size) for g in grids]), dtype=np.float32)
    cap_v = cap_t = 4 * stacked.size
    xyz = np.full(3 * cap_v, np.nan, np.float32)
    idx = np.full(3 * cap_t, 0xFFFFFFFF, np.uint32)
    tot = np.zeros(6, np.uint64)
    cv, ct = np.zeros(batch + 1, np.uint64), np.zeros(batch + 1, np.uint64)
    rc = model.list_model_extract_batch(size, batch, stacked.ctypes.data, n_warps, 3, 1 << 14, xyz.ctypes.data, cap_v, idx.ctypes.data,
                                        cap_t, tot.ctypes.data, cv.ctypes.data, ct.ctypes.data)
    assert rc == 0
    for b in range(batch):
        oxyz, oidx, _ = oracle.extract_grid(size, grids[b])
        v0, v1, t0, t1 = int(cv[b]), int(cv[b + 1]), int(ct[b]), int(ct[b + 1])
        assert (v1 - v0, t1 - t0) == (len(oxyz) // 3, len(oidx) // 3), "chunk %d counts" % b
        assert mesh_diff(xyz[3 * v0:3 * v1], idx[3 * t0:3 * t1], oxyz, oidx) == "", "chunk %d" % b
    assert int(tot[0]) == int(cv[batch]) and int(tot[2]) == int(ct[batch])
