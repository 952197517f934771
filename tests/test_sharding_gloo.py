"""N > 1 host path on CPU: world_size 2 over gloo.  Each rank computes the {V, V_before_last_layer, T}
totals of its z-slab from the oracle's mesh, exchanges them with the same all-gather helper the GPU path
uses, derives its bases, renumbers its slab-local mesh -- the concatenation must equal the global mesh."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _slab_mesh_from_oracle(O, size, grid, z0, z1):
    """What the rank owning cell layers [z0, z1) produces: its own vertices, and its own triangles with
    ids counted from the first vertex of its ghost layer z0-1 (the device numbers a slab from there, using
    the *global* z for the z == 0 ownership extras, so a slab's numbering is a window of the global one)."""
    xyz, idx, _ = O.extract_grid(size, grid)
    cum_v, cum_t = [0], [0]
    for z in range(1, size + 1):  # the mesh of the first z cell layers is an exact prefix (see test_oracle)
        wx, wi, _ = O.extract_grid(size, grid[:z + 1], z_cells=z)
        cum_v.append(len(wx) // 3)
        cum_t.append(len(wi) // 3)
    lo = z0 - (1 if z0 > 0 else 0)
    own_xyz = xyz[3 * cum_v[z0]:3 * cum_v[z1]]
    own_idx_local = idx[3 * cum_t[z0]:3 * cum_t[z1]].astype(np.int64) - cum_v[lo]
    assert own_idx_local.min() >= 0
    totals = [cum_v[z1] - cum_v[z0], cum_v[z1 - 1] - cum_v[z0], cum_t[z1] - cum_t[z0]]
    return own_xyz, own_idx_local, totals


def _worker(rank, world, port, size, q):
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from helpers import oracle_prog
    from isosurface_b200.sharded import allgather_totals, bases_from_totals, slab_range
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = O.fill_grid_sdf(size, oracle_prog("csgA"))
    z0, z1 = slab_range(size, rank, world)
    own_xyz, own_idx_local, totals = _slab_mesh_from_oracle(O, size, grid, z0, z1)
    gathered = allgather_totals(totals)
    vbase, bbase, tbase = bases_from_totals(gathered, rank)
    ofs = bbase if z0 > 0 else vbase                        # what isomc_slab_emit uploads as the id offset
    own_idx = (own_idx_local + ofs).astype(np.uint32)
    q.put((rank, own_xyz, own_idx, vbase, tbase))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_concatenation_is_the_global_mesh(oracle):
    from helpers import oracle_prog
    size, world = 24, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, size, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xyz, idx, _ = oracle.extract_sdf(size, oracle_prog("csgA"))
    cat_xyz = np.concatenate([p[1] for p in parts])
    cat_idx = np.concatenate([p[2] for p in parts])
    assert np.array_equal(cat_idx, idx)
    assert np.array_equal(cat_xyz.view(np.uint32), xyz.view(np.uint32))
    assert parts[1][3] == len(parts[0][1]) // 3 and parts[1][4] == len(parts[0][2]) // 3


def _rebalance_worker(rank, world, port, size, q):
    """the host logic of bench.strong_record / a production host: per-layer active-cell counts of each rank's equal slab are
    summed into one lattice-long vector (all-reduce), every rank cuts the same slabs of equal work from it, and the re-cut
    slabs still concatenate to the global mesh"""
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from helpers import oracle_prog
    from isosurface_b200.sharded import allgather_totals, balanced_slabs, bases_from_totals, slab_range
    from oracle import oracle as O
    dist.init_process_group("gloo", rank=rank, world_size=world)
    grid = O.fill_grid_sdf(size, oracle_prog("sphere05_origin"))  # an octant of a sphere at the origin: the equal split is lopsided
    z0, z1 = slab_range(size, rank, world)
    ci = O.cube_indices(size, grid)
    act = ((ci != 0) & (ci != 255)).reshape(size, -1).sum(axis=1)
    layers = torch.zeros(size, dtype=torch.int64)
    layers[z0:z1] = torch.from_numpy(act[z0:z1].astype(np.int64))   # what isomc_layer_counts gives this rank
    dist.all_reduce(layers)
    slabs = balanced_slabs(size, layers.numpy(), world, active_cell_cost=2000.0)
    b0, b1 = slabs[rank]
    own_xyz, own_idx_local, totals = _slab_mesh_from_oracle(O, size, grid, b0, b1)
    gathered = allgather_totals(totals)
    vbase, bbase, tbase = bases_from_totals(gathered, rank)
    own_idx = (own_idx_local + (bbase if b0 > 0 else vbase)).astype(np.uint32)
    q.put((rank, own_xyz, own_idx, slabs, [int(x) for x in layers.tolist()] == [int(x) for x in act.tolist()]))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_rebalanced_slabs(oracle):
    from helpers import oracle_prog
    from isosurface_b200.sharded import slab_range
    size, world = 24, 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rebalance_worker, args=(r, world, port, size, q)) for r in range(world)]
    for p in procs:
        p.start()
    parts = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert parts[0][3] == parts[1][3] and parts[0][3] != [slab_range(size, r, world) for r in range(world)]  # same cut everywhere, not the equal one
    assert parts[0][4] and parts[1][4]                       # the all-reduced vector is the whole lattice's per-layer count
    xyz, idx, _ = oracle.extract_sdf(size, oracle_prog("sphere05_origin"))
    assert np.array_equal(np.concatenate([p[2] for p in parts]), idx)
    assert np.array_equal(np.concatenate([p[1] for p in parts]).view(np.uint32), xyz.view(np.uint32))
