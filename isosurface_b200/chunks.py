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
