"""z-slab sharding of one extract across the GPUs of a node.

The reference has no multi-chunk driver; this is new.  Rank g of G owns cell layers
[z_g, z_{g+1}) and is handed sample layers [z_g - (g > 0), z_{g+1}].  The only exchange is one
all-gather of three u64 per rank {V_owned, V_owned_before_last_cell_layer, T_owned}; from it every
rank derives the offset that makes its indices global, and the emission kernel writes global ids
directly (no re-index pass).  Concatenating the ranks' outputs in rank order is the reference mesh.

Host logic only; the collective goes through `torch.distributed` (NCCL on GPUs, gloo in CPU tests).
"""
import ctypes as C

import numpy as np

from . import _lib


def slab_range(size, rank, world):
    """cell layers [z0, z1) of `rank`: contiguous, balanced to within one layer"""
    if not (0 <= rank < world) or world > size:
        raise ValueError("bad rank/world %d/%d for size %d" % (rank, world, size))
    base, rem = divmod(size, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


def slab_sample_layers(size, rank, world):
    """(first sample layer, number of sample layers) a rank must hold"""
    z0, z1 = slab_range(size, rank, world)
    ghost = 1 if z0 > 0 else 0
    return z0 - ghost, (z1 - z0) + ghost + 1


# cost of one active cell in units of one lattice sample, measured on a B200 (profiles/r02_history.md: an eighth of the 2048^3
# sphere field costs 0.85 ms + 0.06 ms per million active cells; 0.85 ms / (256 * 2048^2 samples) = 0.79 ps per sample)
ACTIVE_CELL_COST = 76.0


def balanced_slabs(size, layer_active, world, active_cell_cost=ACTIVE_CELL_COST):
    """Contiguous cell-layer ranges [(z0, z1)] * world of (nearly) equal WORK instead of equal thickness.

    layer_active[z] = active cells of cell layer z in an earlier extract of a similar field (isomc_layer_counts; all-gathered
    over the ranks of the earlier split).  Work of a layer = size^2 samples to stream + active_cell_cost per active cell
    (classification, list entry, look-ups, vertices and triangles).  Every rank gets at least one layer; boundaries are the
    points where the running work crosses k / world of the total."""
    a = np.asarray(layer_active, dtype=np.float64)
    if a.shape != (size,) or world < 1 or world > size:
        raise ValueError("need one count per cell layer (%d) and 1 <= world <= size" % size)
    work = float(size) * float(size) + active_cell_cost * a
    cum = np.concatenate([[0.0], np.cumsum(work)])
    cuts = [0]
    for k in range(1, world):
        z = int(np.searchsorted(cum, cum[-1] * k / world, side="left"))
        z = max(z, cuts[-1] + 1)            # at least one layer per rank ...
        z = min(z, size - (world - k))      # ... including the ranks still to come
        cuts.append(z)
    cuts.append(size)
    return [(cuts[k], cuts[k + 1]) for k in range(world)]


def bases_from_totals(gathered, rank):
    """gathered: (G, 3) integer array of per-rank {V, V_before_last_layer, T}.
    Returns (vertex_base, boundary_base, triangle_base) of `rank`:
      vertex_base   = sum of V of lower ranks (id of this rank's first own vertex)
      boundary_base = id of the first vertex created in the previous rank's last cell layer
                      (= vertex_base[rank-1] + V_before_last_layer[rank-1]); 0 for rank 0
      triangle_base = sum of T of lower ranks (position of this rank's triangles in the global list)
    Mirrors the device-side k_slab_bases."""
    g = np.asarray(gathered, dtype=np.uint64).reshape(-1, 3)
    vbase = int(g[:rank, 0].sum())
    tbase = int(g[:rank, 2].sum())
    bbase = 0
    if rank > 0:
        bbase = vbase - int(g[rank - 1, 0]) + int(g[rank - 1, 1])
    return vbase, bbase, tbase


def allgather_totals(totals, group=None):
    """all-gather 3 integers per rank through torch.distributed (any backend) -> (G, 3) uint64 array"""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = "cuda" if backend == "nccl" else "cpu"
    mine = torch.tensor([int(t) for t in totals], dtype=torch.int64, device=dev)
    out = torch.zeros(3 * world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(out, mine, group=group)
    return out.cpu().numpy().astype(np.uint64).reshape(world, 3)


class SlabMarchingCubes:
    """One rank's share of a sharded extract.  `extract(d_slab_ptr)` runs count -> all-gather -> emit.
    z_range: the rank's cell layers (default: the equal split `slab_range`; `balanced_slabs` gives a split of equal work)."""

    def __init__(self, size, rank, world, device=0, z_range=None):
        lib = _lib.load()
        self.size, self.rank, self.world, self.device = int(size), int(rank), int(world), int(device)
        self.z0, self.z1 = slab_range(size, rank, world) if z_range is None else (int(z_range[0]), int(z_range[1]))
        self._h = C.c_void_p()
        _lib.check(lib.isomc_slab_create(self.size, self.z0, self.z1, self.device, C.byref(self._h)))
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

    def count(self, d_slab_ptr):
        _lib.check(self._lib.isomc_slab_count_grid_device(self._h, C.c_void_p(d_slab_ptr)), self._h)
        t = (C.c_uint64 * 3)()
        _lib.check(self._lib.isomc_slab_totals(self._h, C.byref(t)), self._h)
        return [int(t[0]), int(t[1]), int(t[2])]

    def emit(self, vertex_base, boundary_base):
        _lib.check(self._lib.isomc_slab_emit(self._h, int(vertex_base), int(boundary_base)), self._h)

    def extract(self, d_slab_ptr, gathered=None, group=None):
        """gathered: optional precomputed (G,3) totals (single-process simulation of all ranks)"""
        mine = self.count(d_slab_ptr)
        if gathered is None:
            gathered = allgather_totals(mine, group) if self.world > 1 else np.array([mine], dtype=np.uint64)
        vbase, bbase, tbase = bases_from_totals(gathered, self.rank)
        self.emit(vbase, bbase)
        return vbase, tbase

    def copy_out(self):
        v, t, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        _lib.check(self._lib.isomc_counts(self._h, C.byref(v), C.byref(t), C.byref(a)), self._h)
        xyz = np.empty(v.value * 3, np.float32)
        idx = np.empty(t.value * 3, np.uint32)
        _lib.check(self._lib.isomc_copy_out(self._h, xyz.ctypes.data, idx.ctypes.data), self._h)
        return xyz, idx

    def layer_active_cells(self):
        """active cells of each of this rank's cell layers in the last extract (input of `balanced_slabs`)"""
        n = self.z1 - self.z0
        buf = np.zeros(3 * n, np.uint64)
        _lib.check(self._lib.isomc_layer_counts(self._h, buf.ctypes.data), self._h)
        return buf.reshape(n, 3)[:, 2].copy()

    # ---- the exchange as peer stores over NVLink (no collective call per step) ----------------
    def connect_peers(self, group=None):
        """One process per GPU: exchange the CUDA IPC handles of the ranks' mailboxes ONCE (through torch.distributed, any
        backend) and map them; afterwards `extract_exchanged` needs no collective and no host round trip per step."""
        connect_peers_ipc(self._lib, self._h, self.rank, self.world, group)

    def extract_exchanged(self, d_slab_ptr):
        """count -> totals to every rank's mailbox / wait / id offset on the device -> emit"""
        _lib.check(self._lib.isomc_slab_extract_grid_exchanged(self._h, C.c_void_p(d_slab_ptr)), self._h)


def connect_peers_ipc(lib, handle, rank, world, group=None):
    """all-gather the 64-byte CUDA IPC handles of the slab mailboxes and connect `handle` to its peers"""
    import torch
    import torch.distributed as dist
    mine = (C.c_uint8 * 64)()
    _lib.check(lib.isomc_slab_mailbox_ipc(handle, mine), handle)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t_mine = torch.tensor(list(mine), dtype=torch.uint8, device=dev)
    t_all = torch.zeros(64 * world, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(t_all, t_mine, group=group)
    handles = np.ascontiguousarray(t_all.cpu().numpy())
    try:
        _lib.check(lib.isomc_slab_connect_ipc(handle, rank, world, handles.ctypes.data), handle)
    finally:
        dist.barrier(group=group)  # every rank has mapped every mailbox (or given up) before anyone publishes


class ShardedMarchingCubes:
    """`MarchingCubes(size)` over several GPUs of one box, driven from THIS process through the in-library
    `isomc_sharded_*` entry points (include/isomc.h): one slab handle per device, one exchange of 3 x u64 per rank per
    extract (peer stores into mailboxes over NVLink when the devices are peers, else an NCCL all-gather).  `devices` may list one device several times (a single-GPU box exercising the sharded path)."""

    def __init__(self, size, devices):
        lib = _lib.load()
        self.size, self.devices = int(size), [int(d) for d in devices]
        self._h = C.c_void_p()
        arr = (C.c_int32 * len(self.devices))(*self.devices)
        rc = lib.isomc_sharded_create(self.size, len(self.devices), arr, C.byref(self._h))
        if rc:
            raise _lib.IsomcError(rc, (lib.isomc_sharded_last_error(None) or b"").decode())
        self._lib = lib

    def _check(self, rc):
        if rc:
            raise _lib.IsomcError(rc, (self._lib.isomc_sharded_last_error(self._h) or b"").decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.isomc_sharded_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def uses_nccl(self):
        return bool(self._lib.isomc_sharded_uses_nccl(self._h))

    @property
    def uses_peer_memory(self):
        return bool(self._lib.isomc_sharded_uses_peer_memory(self._h))

    def slab(self, rank):
        """(z_begin, z_end, first sample layer, number of sample layers) of `rank`"""
        v = [C.c_uint32() for _ in range(4)]
        self._check(self._lib.isomc_sharded_slab(self._h, rank, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    def extract_grid(self, slab_ptrs):
        """slab_ptrs[r]: device pointer (on devices[r]) to rank r's sample layers"""
        arr = (C.c_void_p * len(self.devices))(*[int(p) for p in slab_ptrs])
        self._check(self._lib.isomc_sharded_extract_grid(self._h, arr))
        return self.counts()

    def extract_sdf(self, source, distance="signed"):
        """distance="directed": MarchingCubes<Directed> over the slabs (the tree sampled through sample_vector)"""
        from .source import Sampler, encode_program
        prog = encode_program(source.source if isinstance(source, Sampler) else source)
        fn = self._lib.isomc_sharded_extract_sdf_directed if distance == "directed" else self._lib.isomc_sharded_extract_sdf
        self._check(fn(self._h, prog.ctypes.data, len(prog)))
        return self.counts()

    def counts(self):
        v, t, a = C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._check(self._lib.isomc_sharded_counts(self._h, C.byref(v), C.byref(t), C.byref(a)))
        return v.value, t.value, a.value

    def copy_out(self):
        nv, nt, _ = self.counts()
        xyz = np.empty(nv * 3, np.float32)
        idx = np.empty(nt * 3, np.uint32)
        self._check(self._lib.isomc_sharded_copy_out(self._h, xyz.ctypes.data, idx.ctypes.data))
        return xyz, idx
