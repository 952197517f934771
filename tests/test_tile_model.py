"""Host model of the tile path (tests/tile_model.cu) against the oracle.

The model executes the kernels' own source (isomc_tile.cuh: tile_count_item, tile_emit_item) on the CPU with
emulated CTAs -- 256 coroutines, emulated warp shuffles and block barriers, work items and threads run in shuffled
order -- so the entry format, the per-piece prefixes, the slot ring, the edge-id planes with their halo row / halo
column / previous-layer protocol and the block allocators are checked bit for bit without a GPU.  The GPU parity
tests then only have to confirm the launch plumbing and the TMA copies.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from helpers import mesh_diff, oracle_prog

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "tile_model.cu"
SO = ROOT / "tests" / "_build" / "libtile_model.so"
DEPS = [SRC] + [ROOT / "isosurface_b200" / "csrc" / n for n in
                ("isomc_tile.cuh", "isomc_cell.cuh", "isomc_device.cuh", "isomc_tables.h", "isomc_case_table.h")]


@pytest.fixture(scope="module")
def model():
    SO.parent.mkdir(exist_ok=True)
    if not SO.exists() or any(d.stat().st_mtime > SO.stat().st_mtime for d in DEPS):
        subprocess.run(["nvcc", "-O2", "-std=c++17", "-arch=sm_100a", "-DISOMC_HOST_MODEL", "-diag-suppress", "177,20011,20014",
                        "-Xcompiler", "-ffp-contract=off,-fPIC,-fno-fast-math", "-shared", "-o", str(SO), str(SRC)], check=True)
    lib = C.CDLL(str(SO))
    lib.tile_model_extract.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                       C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.tile_model_extract_directed.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                                C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
    lib.tile_model_sample_vector.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
    return lib


def run_model(lib, size, grid, z_begin=0, z_end=None, n_ctas=5, seed=1, zc=4, ring=3, vofs=0, cap_eb=4096, cap_tb=4096, cap_v=None,
              cap_t=None):
    z_end = size if z_end is None else z_end
    ghost = 1 if z_begin > 0 else 0
    slab = np.ascontiguousarray(grid[z_begin - ghost:z_end + 1], dtype=np.float32)
    cap_v = 4 * slab.size if cap_v is None else cap_v
    cap_t = 4 * slab.size if cap_t is None else cap_t
    xyz = np.full(3 * cap_v, np.nan, np.float32)
    idx = np.full(3 * cap_t, 0xFFFFFFFF, np.uint32)
    tot = np.zeros(8, np.uint64)
    rc = lib.tile_model_extract(size, z_begin, z_end, slab.ctypes.data, n_ctas, seed, zc, ring, vofs, cap_eb, cap_tb, xyz.ctypes.data,
                                cap_v, idx.ctypes.data, cap_t, tot.ctypes.data)
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


@pytest.mark.parametrize("size,seed,n_ctas,zc,ring", [(2, 1, 1, 4, 3), (3, 2, 3, 1, 2), (17, 3, 5, 3, 3), (33, 4, 9, 16, 3), (34, 5, 2, 5, 2),
                                                      (65, 6, 11, 7, 3), (70, 7, 64, 64, 3)])
def test_model_matches_oracle_on_noise(model, oracle, size, seed, n_ctas, zc, ring):
    """white noise: ~every cell active, every boundary-ownership case, dense rows (32 cells per segment, 5 triangles each);
    several tile rows (halo rows), z-chunks of every length (previous-layer warm-up), both slot rings"""
    grid = noise(size, seed)
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    rc, xyz, idx, tot = run_model(model, size, grid, n_ctas=n_ctas, seed=seed, zc=zc, ring=ring)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


@pytest.mark.parametrize("size,zc_layers,n_ctas", [(520, 2, 3), (1030, 1, 4), (1060, 2, 7)])
def test_model_wide_rows(model, oracle, size, zc_layers, n_ctas):
    """rows of more than one tile (N > 513): row pieces, halo columns, ids relative to a neighbouring piece; a thin z window"""
    rng = np.random.default_rng(size)
    grid = rng.standard_normal((zc_layers + 1, size, size)).astype(np.float32)
    grid[:, :, 200:480] = np.abs(grid[:, :, 200:480])           # long empty stretches
    grid[:, 300:310, 100:size - 1] = -np.abs(grid[:, 300:310, 100:size - 1])
    grid[:, ::7, :] = 1.0                                       # empty rows
    oxyz, oidx, oact = oracle.extract_grid(size, grid, z_cells=zc_layers)
    rc, xyz, idx, tot = run_model(model, size, grid, z_end=zc_layers, n_ctas=n_ctas, zc=1, cap_eb=40000, cap_tb=40000)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


def test_model_sparse_field(model, oracle):
    """a sparse field: most tile layers are empty (skipped), the surface enters and leaves tiles"""
    size = 96
    prog = oracle_prog("sphere03")
    grid = oracle.fill_grid_sdf(size, prog)
    grid[40:44, 10:12, 3:90] *= -1.0  # some scattered sign flips
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    for n_ctas in (1, 4, 29):
        rc, xyz, idx, tot = run_model(model, size, grid, n_ctas=n_ctas, seed=n_ctas, zc=8)
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


def test_model_overflow_is_reported(model, oracle):
    size = 33
    grid = noise(size, 5)
    _, _, oact = oracle.extract_grid(size, grid)
    rc, _, _, tot = run_model(model, size, grid, cap_eb=8, cap_tb=8)
    assert rc == 1 and tot[3] == oact and tot[4] > 8 and tot[5] > 8  # totals stay right, the host can grow and re-run
    rc, xyz, idx, tot2 = run_model(model, size, grid, cap_eb=tot[4], cap_tb=tot[5])
    assert rc == 0 and tot2[:4] == tot[:4]


def test_model_special_values(model, oracle):
    """+-0, +-inf and NaN samples: `!(v > 0)` classification and crossing parameters as the reference computes them"""
    size = 20
    grid = noise(size, 9)
    flat = grid.reshape(-1)
    rng = np.random.default_rng(3)
    for v in (0.0, -0.0, np.inf, -np.inf, np.nan):
        flat[rng.integers(0, flat.size, 200)] = v
    oxyz, oidx, oact = oracle.extract_grid(size, grid)
    rc, xyz, idx, tot = run_model(model, size, grid, zc=3)
    assert rc == 0 and tot[3] == oact
    assert mesh_diff(xyz, idx, oxyz, oidx) == ""


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
    assert model.tile_model_sample_vector(prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), got.ctypes.data) == 0
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), "%d components differ, first at point %s" % (int((~same).sum()), pts[np.argwhere(~same)[0][0]])


@pytest.mark.parametrize("name,size", [("sphere03", 32), ("torus", 40), ("csgA", 48), ("csgB", 33), ("torus_origin", 64), ("nested", 36)])
def test_model_directed_extract_matches_oracle(model, oracle, name, size):
    """MarchingCubes<Directed>: the tile kernels' source with the three-component source against the restated reference"""
    prog = oracle_prog(name)
    oxyz, oidx, oact = oracle.extract_sdf_directed(size, prog)
    cap_v, cap_t = len(oxyz) // 3 + 64, len(oidx) // 3 + 64
    xyz = np.full(3 * cap_v, np.nan, np.float32)
    idx = np.full(3 * cap_t, 0xFFFFFFFF, np.uint32)
    tot = np.zeros(8, np.uint64)
    rc = model.tile_model_extract_directed(size, prog.ctypes.data, len(prog), 5, 3, 6, 4096, 4096, xyz.ctypes.data, cap_v, idx.ctypes.data,
                                           cap_t, tot.ctypes.data)
    assert rc == 0 and int(tot[3]) == oact
    assert mesh_diff(xyz[:3 * int(tot[0])], idx[:3 * int(tot[2])], oxyz, oidx) == ""
