"""ctypes loader for the CPU oracle (oracle/mc_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this module.  The product package (isosurface_b200/) never does.
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

# op codes: same numbering as include/isomc.h (ISOMC_SDF_*), restated independently
SPHERE, TORUS, CYLINDER, PRISM = 1, 2, 3, 4
UNION, INTERSECTION, DIFFERENCE = 16, 17, 18
TRANSLATE_PUSH, TRANSLATE_POP = 32, 33
FAITHFUL, LEAN = 0, 1

NODE_DTYPE = np.dtype([("op", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4")])


class _Mesh(C.Structure):
    _fields_ = [("xyz", C.POINTER(C.c_float)), ("n_vertices", C.c_uint64), ("cap_v", C.c_uint64),
                ("idx", C.POINTER(C.c_uint32)), ("n_triangles", C.c_uint64), ("cap_i", C.c_uint64),
                ("n_active_cells", C.c_uint64)]


def build():
    subprocess.run(["make", "-s", "-C", str(_HERE)], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = _HERE / "libmc_oracle.so"
        if not so.exists() or so.stat().st_mtime < max((_HERE / n).stat().st_mtime for n in ("mc_oracle.c", "synth_field.c", "Makefile")):
            build()
        _LIB = C.CDLL(str(so))
        _LIB.oracle_extract_sdf.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(_Mesh)]
        _LIB.oracle_extract_grid.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_int, C.POINTER(_Mesh)]
        _LIB.oracle_cube_indices.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        _LIB.oracle_fill_grid_sdf.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
        _LIB.oracle_sample_sdf.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        _LIB.oracle_tables.argtypes = [C.c_void_p] * 4
        _LIB.oracle_mesh_free.argtypes = [C.POINTER(_Mesh)]
        _LIB.oracle_interleaved_normals_cd.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.c_void_p, C.c_uint32, C.c_void_p,
                                                       C.c_uint64, C.c_void_p]
        _LIB.oracle_sample_sdf_vector.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p]
        _LIB.oracle_extract_sdf_directed.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(_Mesh)]
        _LIB.oracle_point_cloud_sdf.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(_Mesh)]
        _LIB.oracle_point_cloud_sdf_directed.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(_Mesh)]
        _LIB.oracle_point_cloud_grid.argtypes = [C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(_Mesh)]
        _LIB.oracle_synth_field.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p]
    return _LIB


def program(nodes):
    """nodes: iterable of (op, a, b, c) -> packed node array."""
    arr = np.zeros(len(nodes), dtype=NODE_DTYPE)
    for i, nd in enumerate(nodes):
        nd = tuple(nd) + (0.0,) * (4 - len(nd))
        arr[i] = nd
    return arr


def _take(mesh):
    nv, nt = int(mesh.n_vertices), int(mesh.n_triangles)
    xyz = np.ctypeslib.as_array(mesh.xyz, shape=(nv * 3,)).copy() if nv else np.zeros(0, np.float32)
    idx = np.ctypeslib.as_array(mesh.idx, shape=(nt * 3,)).copy() if nt else np.zeros(0, np.uint32)
    act = int(mesh.n_active_cells)
    lib().oracle_mesh_free(C.byref(mesh))
    return xyz.astype(np.float32, copy=False), idx.astype(np.uint32, copy=False), act


def extract_sdf(size, prog, mode=LEAN):
    m = _Mesh()
    prog = np.ascontiguousarray(prog)
    rc = lib().oracle_extract_sdf(size, prog.ctypes.data, len(prog), mode, C.byref(m))
    if rc:
        raise RuntimeError("oracle_extract_sdf rc=%d" % rc)
    return _take(m)


def extract_grid(size, grid, z_cells=None, mode=LEAN):
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    z_cells = size if z_cells is None else z_cells
    assert grid.size >= size * size * (z_cells + 1)
    m = _Mesh()
    rc = lib().oracle_extract_grid(size, grid.ctypes.data, z_cells, mode, C.byref(m))
    if rc:
        raise RuntimeError("oracle_extract_grid rc=%d" % rc)
    return _take(m)


def sample_sdf_vector(prog, pts):
    """VectorSource::sample_vector of the implicit tree at pts -> (n, 3) Directed distances"""
    prog = np.ascontiguousarray(prog)
    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((len(pts), 3), dtype=np.float32)
    rc = lib().oracle_sample_sdf_vector(prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_sample_sdf_vector rc=%d" % rc)
    return out


def extract_sdf_directed(size, prog):
    """MarchingCubes::<Directed>::new(size).extract (reference src/distance.rs:72-104)"""
    m = _Mesh()
    prog = np.ascontiguousarray(prog)
    rc = lib().oracle_extract_sdf_directed(size, prog.ctypes.data, len(prog), C.byref(m))
    if rc:
        raise RuntimeError("oracle_extract_sdf_directed rc=%d" % rc)
    return _take(m)


def interleaved_normals_cd(inner_prog, xyz, epsilon=0.000001, offsets=()):
    """IndexedInterleavedNormals over CentralDifference(inner_prog) (reference src/extractor.rs:113-122, src/source.rs:82-94),
    the vertex first moved by `offsets` (q = p - o, outermost first) as DemoSource does: (V, 6) floats"""
    inner_prog = np.ascontiguousarray(inner_prog)
    xyz = np.ascontiguousarray(xyz, dtype=np.float32).reshape(-1, 3)
    off = np.ascontiguousarray(np.asarray(offsets, np.float32).reshape(-1, 3))
    out = np.zeros((len(xyz), 6), np.float32)
    rc = lib().oracle_interleaved_normals_cd(inner_prog.ctypes.data, len(inner_prog), epsilon, off.ctypes.data, len(off),
                                             xyz.ctypes.data, len(xyz), out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_interleaved_normals_cd rc=%d" % rc)
    return out


def point_cloud_sdf(size, prog):
    """PointCloud::<Signed>::new(size).extract (reference src/point_cloud.rs:50-63): xyz of one point per active cell"""
    m = _Mesh()
    prog = np.ascontiguousarray(prog)
    rc = lib().oracle_point_cloud_sdf(size, prog.ctypes.data, len(prog), C.byref(m))
    if rc:
        raise RuntimeError("oracle_point_cloud_sdf rc=%d" % rc)
    return _take(m)[0]


def point_cloud_sdf_directed(size, prog):
    """PointCloud::<Directed>::new(size).extract over an implicit tree (point_cloud.rs:50-63, distance.rs:77-80)"""
    m = _Mesh()
    prog = np.ascontiguousarray(prog)
    rc = lib().oracle_point_cloud_sdf_directed(size, prog.ctypes.data, len(prog), C.byref(m))
    if rc:
        raise RuntimeError("oracle_point_cloud_sdf_directed rc=%d" % rc)
    return _take(m)[0]


def point_cloud_grid(size, grid, z_cells=None):
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    z_cells = size if z_cells is None else z_cells
    m = _Mesh()
    rc = lib().oracle_point_cloud_grid(size, grid.ctypes.data, z_cells, C.byref(m))
    if rc:
        raise RuntimeError("oracle_point_cloud_grid rc=%d" % rc)
    return _take(m)[0]


def cube_indices(size, grid, z_cells=None):
    grid = np.ascontiguousarray(grid, dtype=np.float32)
    z_cells = size if z_cells is None else z_cells
    out = np.zeros((z_cells, size - 1, size - 1), dtype=np.uint8)
    rc = lib().oracle_cube_indices(size, grid.ctypes.data, z_cells, out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_cube_indices rc=%d" % rc)
    return out


def fill_grid_sdf(size, prog, z_layers=None):
    z_layers = size + 1 if z_layers is None else z_layers
    prog = np.ascontiguousarray(prog)
    out = np.zeros((z_layers, size, size), dtype=np.float32)
    rc = lib().oracle_fill_grid_sdf(size, prog.ctypes.data, len(prog), z_layers, out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_fill_grid_sdf rc=%d" % rc)
    return out


def sample_sdf(prog, pts):
    prog = np.ascontiguousarray(prog)
    pts = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
    out = np.zeros(len(pts), dtype=np.float32)
    rc = lib().oracle_sample_sdf(prog.ctypes.data, len(prog), pts.ctypes.data, len(pts), out.ctypes.data)
    if rc:
        raise RuntimeError("oracle_sample_sdf rc=%d" % rc)
    return out


def tables():
    tri = np.zeros((256, 16), np.int8)
    em = np.zeros(256, np.uint16)
    co = np.zeros((8, 3), np.int32)
    ed = np.zeros((12, 2), np.int32)
    rc = lib().oracle_tables(tri.ctypes.data, em.ctypes.data, co.ctypes.data, ed.ctypes.data)
    if rc:
        raise RuntimeError("oracle_tables rc=%d" % rc)
    return tri, em, co, ed


FIELD_FBM, FIELD_GYROID, FIELD_SPHERE_UNION = 1, 2, 3


def synth_field(kind, size, seed, z_first=0, n_layers=None):
    """host-generated benchmark field, shape (n_layers, size, size); see oracle/synth_field.c"""
    n_layers = size + 1 - z_first if n_layers is None else n_layers
    out = np.empty((n_layers, size, size), np.float32)
    rc = lib().oracle_synth_field(kind, size, seed, z_first, n_layers, out.ctypes.data)
    if rc:
        raise ValueError("oracle_synth_field failed (%d)" % rc)
    return out
