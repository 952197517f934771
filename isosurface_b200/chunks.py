"""Multi-chunk driver: many small `size^3` extracts in flight at once.

The reference's usage model is one `MarchingCubes::new(size)` per chunk size and one `extract` per chunk
(reference src/marching_cubes.rs:44-45, README.md:19), the caller moving the source into the unit cube of each chunk.
A 32^3 or 128^3 extract is launch-latency bound on a B200 (four launches and one synchronisation, 60-70 us for a few
microseconds of work), so the way to chunk throughput is overlap, not a faster kernel: every handle owns a CUDA stream
(`&mut self` semantics per handle, distinct handles independent), and `isomc_enqueue_sdf` / `isomc_finish` split an
extract into "launch everything" and "wait + size + deliver".  This driver keeps `n_inflight` handles busy round robin.
Results are the per-chunk meshes of the plain API, bit for bit, in submission order.
"""
from .marching_cubes import MarchingCubes


class ChunkedMarchingCubes:
    """`ChunkedMarchingCubes(size, n_inflight).extract_many(sources)` -> [(xyz, idx), ...]

    sources: implicit sources (or device-resident DenseGrids) of the chunks; each is extracted exactly as
    `MarchingCubes(size).extract_device(source)` + `copy_out()` would."""

    def __init__(self, size, n_inflight=8, device=0):
        if n_inflight < 1:
            raise ValueError("n_inflight must be >= 1")
        self.size = int(size)
        self._pool = [MarchingCubes(size, device=device) for _ in range(int(n_inflight))]

    def close(self):
        for mc in self._pool:
            mc.close()
        self._pool = []

    def extract_many(self, sources, deliver=None):
        """Extract every chunk; `deliver(i, xyz, idx)` is called per chunk in submission order (default: collect a list)."""
        out = [] if deliver is None else None
        pool, k = self._pool, len(self._pool)
        pending = []  # (chunk index, handle) in submission order
        for i, src in enumerate(sources):
            if len(pending) == k:  # the oldest extract frees its handle
                self._finish(pending.pop(0), deliver, out)
            mc = pool[i % k]
            mc.enqueue(src)
            pending.append((i, mc))
        while pending:
            self._finish(pending.pop(0), deliver, out)
        return out

    @staticmethod
    def _finish(item, deliver, out):
        i, mc = item
        mc.finish()
        xyz, idx = mc.copy_out()
        if deliver is None:
            out.append((xyz, idx))
        else:
            deliver(i, xyz, idx)


class BatchedMarchingCubes:
    """`BatchedMarchingCubes(size, n_chunks).extract_many(sources)` -> [(xyz, idx), ...]

    Up to `n_chunks` implicit sources per call go through ONE sign / count / scan / emit kernel sequence on the device
    (`isomc_extract_sdf_batch`: the chunks' lattices stacked in z, one size read-back, one copy-out), which is what amortises
    the launch and synchronisation latency of a 32^3-sized extract.  Each chunk's mesh is byte for byte what
    `MarchingCubes(size).extract_device(source)` + `copy_out()` returns (indices relative to the chunk's own first vertex)."""

    def __init__(self, size, n_chunks=64, device=0, distance="signed"):
        import ctypes as C
        from . import _lib
        if n_chunks < 1:
            raise ValueError("n_chunks must be >= 1")
        if distance not in ("signed", "directed"):
            raise ValueError("distance must be 'signed' or 'directed'")
        self.distance = distance  # 'directed' = MarchingCubes<Directed> per chunk (implicit sources; a lattice holds scalars)
        self.size, self.n_chunks = int(size), int(n_chunks)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        _lib.check(self._lib.isomc_batch_create(self.size, self.n_chunks, int(device), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.isomc_destroy(self._h)
            self._h.value = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def extract_batch(self, sources):
        """one batch (len(sources) <= n_chunks): returns (xyz, idx, v_offsets, t_offsets) -- the chunks' meshes back to back"""
        import ctypes as C
        import numpy as np
        from . import _lib
        from .source import Sampler, encode_program
        progs = [encode_program(s.source if isinstance(s, Sampler) else s) for s in sources]
        if not 1 <= len(progs) <= self.n_chunks:
            raise ValueError("a batch holds 1..%d chunks, got %d" % (self.n_chunks, len(progs)))
        flat = np.concatenate(progs)
        n_nodes = np.asarray([len(p) for p in progs], np.uint32)
        entry = self._lib.isomc_extract_sdf_batch_directed if self.distance == "directed" else self._lib.isomc_extract_sdf_batch
        _lib.check(entry(self._h, flat.ctypes.data, n_nodes.ctypes.data, len(progs)), self._h)
        return self._results(len(progs))

    def _results(self, n):
        import ctypes as C
        import numpy as np
        from . import _lib
        v, t, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.isomc_counts(self._h, C.byref(v), C.byref(t), C.byref(a)), self._h)
        xyz, idx = np.empty(3 * v.value, np.float32), np.empty(3 * t.value, np.uint32)
        _lib.check(self._lib.isomc_copy_out(self._h, xyz.ctypes.data, idx.ctypes.data), self._h)
        vo, to = np.zeros(self.n_chunks + 1, np.uint64), np.zeros(self.n_chunks + 1, np.uint64)
        _lib.check(self._lib.isomc_batch_offsets(self._h, vo.ctypes.data, to.ctypes.data), self._h)
        return xyz, idx, vo[:n + 1], to[:n + 1]

    def extract_grids(self, lattices):
        """Dense chunks: `lattices` = a float32 array of shape (n, size + 1, size, size) on the host (n <= n_chunks), or a CUDA
        tensor of exactly n_chunks lattices (used in place).  Returns (xyz, idx, v_offsets, t_offsets) like extract_batch."""
        if self.distance == "directed":
            raise TypeError("a lattice of scalars has no Directed distances")
        import numpy as np
        from . import _lib
        per = self.size * self.size * (self.size + 1)
        if hasattr(lattices, "data_ptr"):
            if not lattices.is_cuda or lattices.numel() != per * self.n_chunks or str(lattices.dtype) != "torch.float32" or not lattices.is_contiguous():
                raise ValueError("a device batch is a contiguous float32 CUDA tensor of exactly n_chunks lattices")
            n = self.n_chunks
            _lib.check(self._lib.isomc_extract_grid_batch_device(self._h, int(lattices.data_ptr()), n), self._h)
        else:
            arr = np.ascontiguousarray(lattices, dtype=np.float32)
            if arr.size % per or not 1 <= arr.size // per <= self.n_chunks:
                raise ValueError("host lattices must be (n, size + 1, size, size) with 1 <= n <= n_chunks")
            n = arr.size // per
            _lib.check(self._lib.isomc_extract_grid_batch_host(self._h, arr.ctypes.data, n), self._h)
        return self._results(n)

    def extract_many(self, sources, deliver=None):
        """any number of chunks, `n_chunks` per kernel sequence; `deliver(i, xyz, idx)` per chunk in submission order"""
        sources = list(sources)
        out = [] if deliver is None else None
        for b0 in range(0, len(sources), self.n_chunks):
            part = sources[b0:b0 + self.n_chunks]
            xyz, idx, vo, to = self.extract_batch(part)
            for j in range(len(part)):
                cx, ci = xyz[3 * int(vo[j]):3 * int(vo[j + 1])], idx[3 * int(to[j]):3 * int(to[j + 1])]
                if deliver is None:
                    out.append((cx, ci))
                else:
                    deliver(b0 + j, cx, ci)
        return out
