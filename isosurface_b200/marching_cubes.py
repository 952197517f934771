"""MarchingCubes: host-side mirror of reference src/marching_cubes.rs:38-82 over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from .extractor import replay
from .source import DenseGrid, Sampler, encode_program


class MarchingCubes:
    """`MarchingCubes(size).extract(source, extractor)` == reference
    `MarchingCubes::<Signed>::new(size).extract(&source, &mut extractor)`.

    `size` lattice points per axis in x and y; the reference's z loop runs one layer further
    (primal_grid.rs:59), so N x N x (N+1) samples and (N-1)^2 x N cells.  One extract at a time
    per instance (`&mut self`).
    """

    def __init__(self, size, device=0, distance="signed"):
        """distance: "signed" = MarchingCubes::<Signed> (scalar distances), "directed" = MarchingCubes::<Directed> (a signed
        distance along each axis, reference src/distance.rs:43-45,72-104; implicit sources only)"""
        if distance not in ("signed", "directed"):
            raise ValueError("distance must be 'signed' or 'directed'")
        lib = _lib.load()
        self.distance = distance
        self.size, self.device = int(size), int(device)
        self._h = C.c_void_p()
        _lib.check(lib.isomc_create(self.size, self.device, C.byref(self._h)))
        self._lib = lib

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.isomc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference API -------------------------------------------------------------------
    def extract(self, source, extractor):
        src = source.source if isinstance(source, Sampler) else source
        if hasattr(extractor, "_bulk_normals"):  # IndexedInterleavedNormals: normals sampled on the device
            self.extract_device(source)
            xyzn, idx = self.copy_out_interleaved_normals(extractor.source, extractor.central_difference.epsilon,
                                                          extractor.outer_translations)
            extractor._bulk_normals(xyzn, idx)
            return
        if isinstance(src, DenseGrid) and not src.on_device:
            xyz, idx = self.extract_host(src)  # host lattice in, host mesh out: one pipelined call
            # extract_host() hands out views of this instance's persistent buffers; an extractor may keep what it is given
            # (ArrayMesh does), so it gets its own copy -- the next extract() must not change an earlier mesh
            xyz, idx = xyz.copy(), idx.copy()
        else:
            self.extract_device(source)
            xyz, idx = self.copy_out()
        replay(extractor, xyz, idx)

    def extract_host(self, grid, xyz=None, idx=None):
        """Host lattice -> host mesh through `isomc_extract_grid_host_to` (copy-in, kernels and copy-out overlap in
        z-chunks).  `xyz` / `idx` are caller buffers (float32 / uint32, ideally pinned); without them the instance keeps
        its own, sized from the previous extract.  Returns VIEWS of the filled parts: with the instance's own buffers they
        are overwritten by the next extract_host() call on this instance."""
        if grid.size != self.size:
            raise ValueError("grid is for size %d, MarchingCubes for %d" % (grid.size, self.size))
        if self.distance == "directed":
            raise TypeError("a dense scalar lattice has no Directed distances; use an implicit source")
        own = xyz is None or idx is None
        if own:
            xyz = getattr(self, "_hxyz", None)
            idx = getattr(self, "_hidx", None)
            if xyz is None:
                xyz = self._hxyz = np.empty(3 * 1024, np.float32)
                idx = self._hidx = np.empty(3 * 1024, np.uint32)
        rc = self._lib.isomc_extract_grid_host_to(self._h, grid.ptr, xyz.ctypes.data, xyz.size // 3, idx.ctypes.data, idx.size // 3)
        if rc == _lib.ERR_BUFFER_TOO_SMALL and own:  # result is on the device: grow (with headroom) and fetch it
            nv, nt, _ = self.counts()
            xyz = self._hxyz = np.empty(3 * (nv + nv // 8 + 1024), np.float32)
            idx = self._hidx = np.empty(3 * (nt + nt // 8 + 1024), np.uint32)
            rc = self._lib.isomc_copy_out(self._h, xyz.ctypes.data, idx.ctypes.data)
        _lib.check(rc, self._h)
        nv, nt, _ = self.counts()
        return xyz[:3 * nv], idx[:3 * nt]

    # ---- device-resident variants ------------------------------------------------------------
    def extract_device(self, source):
        """Run the extraction and leave the mesh in device memory (see `device_buffers`)."""
        src = source.source if isinstance(source, Sampler) else source
        if isinstance(src, DenseGrid):
            if src.size != self.size:
                raise ValueError("grid is for size %d, MarchingCubes for %d" % (src.size, self.size))
            if self.distance == "directed":
                raise TypeError("a dense scalar lattice has no Directed distances; use an implicit source")
            fn = self._lib.isomc_extract_grid_device if src.on_device else self._lib.isomc_extract_grid_host
            _lib.check(fn(self._h, src.ptr), self._h)
        else:
            prog = encode_program(src)
            fn = self._lib.isomc_extract_sdf_directed if self.distance == "directed" else self._lib.isomc_extract_sdf
            _lib.check(fn(self._h, prog.ctypes.data, len(prog)), self._h)
        return self.counts()

    def enqueue(self, source):
        """Enqueue an extract on the handle's stream without synchronising; pair with `finish()`."""
        src = source.source if isinstance(source, Sampler) else source
        if isinstance(src, DenseGrid):
            if not src.on_device:
                raise ValueError("enqueue() needs a device-resident grid")
            if src.size != self.size:
                raise ValueError("grid is for size %d, MarchingCubes for %d" % (src.size, self.size))
            if self.distance == "directed":
                raise TypeError("a dense scalar lattice has no Directed distances; use an implicit source")
            _lib.check(self._lib.isomc_enqueue_grid_device(self._h, src.ptr), self._h)
        else:
            prog = encode_program(src)
            fn = self._lib.isomc_enqueue_sdf_directed if self.distance == "directed" else self._lib.isomc_enqueue_sdf
            _lib.check(fn(self._h, prog.ctypes.data, len(prog)), self._h)

    def finish(self):
        _lib.check(self._lib.isomc_finish(self._h), self._h)
        return self.counts()

    def counts(self):
        v, t, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.isomc_counts(self._h, C.byref(v), C.byref(t), C.byref(a)), self._h)
        return v.value, t.value, a.value

    def copy_out(self):
        nv, nt, _ = self.counts()
        xyz = np.empty(nv * 3, np.float32)
        idx = np.empty(nt * 3, np.uint32)
        _lib.check(self._lib.isomc_copy_out(self._h, xyz.ctypes.data, idx.ctypes.data), self._h)
        return xyz, idx

    def copy_out_interleaved_normals(self, normal_source, epsilon=0.000001, outer_translations=None):
        """(6V,) float32 x y z nx ny nz and (3T,) uint32 of the last extract; normals = central differences of
        `normal_source` (reference src/extractor.rs:113-122, src/source.rs:82-94), evaluated on the device.
        outer_translations: how many enclosing Translate wrappers lie outside the CentralDifference adaptor (None: all)"""
        nv, nt, _ = self.counts()
        prog = encode_program(normal_source)
        xyzn = np.empty(nv * 6, np.float32)
        idx = np.empty(nt * 3, np.uint32)
        outer = 0xFFFFFFFF if outer_translations is None else int(outer_translations)
        _lib.check(self._lib.isomc_copy_out_interleaved_normals_at(self._h, prog.ctypes.data, len(prog), epsilon, outer,
                                                                   xyzn.ctypes.data, idx.ctypes.data), self._h)
        return xyzn, idx

    def device_buffers(self):
        dx, di = C.c_void_p(), C.c_void_p()
        _lib.check(self._lib.isomc_device_buffers(self._h, C.byref(dx), C.byref(di)), self._h)
        return dx.value, di.value

    def reserve(self, n_vertices, n_triangles):
        _lib.check(self._lib.isomc_reserve(self._h, int(n_vertices), int(n_triangles)), self._h)

    def set_stream(self, cuda_stream_ptr):
        _lib.check(self._lib.isomc_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)), self._h)

    def stream(self):
        s = C.c_void_p()
        _lib.check(self._lib.isomc_get_stream(self._h, C.byref(s)), self._h)
        return s.value or 0

    def set_profiling(self, on=True):
        _lib.check(self._lib.isomc_set_profiling(self._h, 1 if on else 0), self._h)

    def stats(self):
        s = _lib.Stats()
        _lib.check(self._lib.isomc_stats_get(self._h, C.byref(s)), self._h)
        return {name: getattr(s, name) for name, _ in s._fields_}

    def cube_indices(self):
        """per-cell cube_index (reference corner order) of the last extract, shape (N, N-1, N-1)"""
        n = self.size
        out = np.zeros((n, n - 1, n - 1), np.uint8)
        _lib.check(self._lib.isomc_debug_cube_indices(self._h, out.ctypes.data), self._h)
        return out


class PointCloud(MarchingCubes):
    """`PointCloud(size).extract(source, extractor)` == reference `PointCloud::<Signed>::new(size).extract(&source,
    &mut extractor)` (src/point_cloud.rs:33-63): one vertex per active cell, the midpoint of the cell's corners 0 and 6,
    in cell order; no face data.  Shares the handle machinery of MarchingCubes (same lattice, same sources)."""

    def extract(self, source, extractor):
        self.extract_device(source)
        xyz, _ = self.copy_out()
        replay(extractor, xyz, np.zeros(0, np.uint32))

    def extract_device(self, source):
        src = source.source if isinstance(source, Sampler) else source
        if isinstance(src, DenseGrid):
            if src.size != self.size:
                raise ValueError("grid is for size %d, PointCloud for %d" % (src.size, self.size))
            if self.distance == "directed":
                raise TypeError("a dense scalar lattice has no Directed distances; use an implicit source")
            fn = self._lib.isomc_points_grid_device if src.on_device else self._lib.isomc_points_grid_host
            _lib.check(fn(self._h, src.ptr), self._h)
        else:
            prog = encode_program(src)
            fn = self._lib.isomc_points_sdf_directed if self.distance == "directed" else self._lib.isomc_points_sdf
            _lib.check(fn(self._h, prog.ctypes.data, len(prog)), self._h)
        return self.counts()

    def extract_host(self, grid, xyz=None, idx=None):
        raise NotImplementedError("PointCloud delivers through extract() / extract_device() + copy_out()")

    def enqueue(self, source):
        raise NotImplementedError("PointCloud extracts synchronously")
