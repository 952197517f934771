#!/usr/bin/env python3
"""bench.py -- MarchingCubes extract throughput on B200 (Gvoxels/s, Mtris/s, % of HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

A "step" is one full `extract` of the workload's field: device-resident dense f32 lattice (or an
implicit SDF program) in, device-resident globally-indexed mesh (xyz f32 + u32 indices) out.

Workloads (BASELINE.json configs / SURVEY.md 8d):
    fbm512       512^3 dense f32 grid, random-phase fBm (C3) -- the default at --gpus 1 (headline)
    fbmweak      default at --gpus N > 1: the same fBm family with 512^3 voxels PER GPU (size 644 / 812 / 1024
                 for N = 2 / 4 / 8, same surface density), z-slab sharded: weak scaling
    gyroid1024   1024^3 dense f32 grid, gyroid (C4)
    spheres2048  2048^3 dense f32 grid, union of 64 spheres (C5); `--workload spheres2048 --gpus N` is the
                 strong-scaling sweep of the north star (recorded in profiles/r01_scale_spheres2048.json)
    torus256 / csga256 / csgb256   256^3 implicit SDF evaluated on device (C2)
    sphere32 / torus128            the reference's CPU-sized cases (C1a / C1b)

N > 1 is launched by torchrun (one rank per GPU); ranks own z-slabs, exchange 3 x u64 totals with one
NCCL all-gather on the extraction stream and write globally numbered indices directly.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (a C restatement of the
reference algorithm; the Rust reference cannot be built in this image) on a bounded sample of the same
workload -- the only place besides tests/ and smoke() where oracle/ is executed.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (size, kind, field/source, seed)
    "fbm512": (512, "grid", "fbm", 0x1505F00D),
    "fbm256": (256, "grid", "fbm", 0x1505F00D),
    "fbm644": (644, "grid", "fbm", 0x1505F00D),    # 2 x 512^3 voxels
    "fbm812": (812, "grid", "fbm", 0x1505F00D),    # 4 x 512^3 voxels
    "fbm1024": (1024, "grid", "fbm", 0x1505F00D),  # 8 x 512^3 voxels
    "gyroid1024": (1024, "grid", "gyroid", 0),
    "gyroid512": (512, "grid", "gyroid", 0),
    "spheres2048": (2048, "grid", "spheres", 0x5EEDBA11),
    "spheres1024": (1024, "grid", "spheres", 0x5EEDBA11),
    "spheres512": (512, "grid", "spheres", 0x5EEDBA11),
    "torus256": (256, "sdf", "torus", 0),
    "csga256": (256, "sdf", "csgA", 0),
    "csgb256": (256, "sdf", "csgB", 0),
    "sphere32": (32, "sdf", "sphere03", 0),
    "torus128": (128, "sdf", "torus_origin", 0),
}
FIELD_KIND = {"fbm": 1, "gyroid": 2, "spheres": 3}
# f32 ops per sample of the implicit shapes, incl. 6 for the lattice coordinate
SDF_OPS = {"torus": 19 + 3, "csgA": 48, "csgB": 30, "sphere03": 17, "torus_origin": 19}


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """samples SM clock / throttle reasons during the timed region (pynvml; nvidia-smi equivalent)"""

    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        self.max_mhz = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join(timeout=1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def slab_range(size, rank, world):
    """contiguous cell-layer ranges balanced to +-1 layer"""
    base, rem = divmod(size, world)
    z0 = rank * base + min(rank, rem)
    return z0, z0 + base + (1 if rank < rem else 0)


class CudaArray:
    """zero-copy torch view of a raw device pointer via __cuda_array_interface__"""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def make_field(lib, torch, dev, wl, z_first, n_layers):
    size, kind, field, seed = WORKLOADS[wl]
    t = torch.empty(n_layers * size * size, dtype=torch.float32, device="cuda:%d" % dev)
    from isosurface_b200 import _lib
    _lib.check(lib.isomc_synth_field(dev, FIELD_KIND[field], size, seed, z_first, n_layers, C.c_void_p(t.data_ptr())))
    return t


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle (C restatement of the reference), faithful-cost mode, 1 thread
# ------------------------------------------------------------------------------------------------

def cpu_baseline(wl, host_layers=None, target_s=12.0):
    """times the oracle on a bounded sample of the workload; returns (dict, mesh_counts)"""
    from oracle import oracle as O
    sys.path.insert(0, str(ROOT / "tests"))
    size, kind, field, seed = WORKLOADS[wl]
    if kind == "sdf":
        from helpers import oracle_prog
        prog = oracle_prog(field)
        t0 = time.perf_counter()
        xyz, idx, act = O.extract_sdf(size, prog, O.FAITHFUL)
        dt = time.perf_counter() - t0
        vox = float(size) ** 3
        return {"value": vox / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
                "sample": "full %s extract (%d^3), oracle faithful-cost mode, 1 thread, %.2f s" % (wl, size, dt),
                "mtris_per_s": len(idx) / 3 / dt / 1e6}
    # grid: first Z cell layers of the same field bytes
    zmax = host_layers.shape[0] - 1
    z = min(4, zmax)
    t0 = time.perf_counter()
    O.extract_grid(size, host_layers, z, O.FAITHFUL)
    dt = time.perf_counter() - t0
    z2 = int(max(z, min(zmax, z * target_s / max(dt, 1e-3))))
    t0 = time.perf_counter()
    xyz, idx, act = O.extract_grid(size, host_layers, z2, O.FAITHFUL)
    dt = time.perf_counter() - t0
    vox = float(size) * size * z2
    return {"value": vox / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
            "sample": "first %d of %d cell layers of %s (same field bytes), oracle faithful-cost mode, 1 thread, %.2f s; "
                      "C restatement of the reference, not cargo bench" % (z2, size, wl, dt),
            "mtris_per_s": len(idx) / 3 / dt / 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    size, kind, field, seed = WORKLOADS[wl]
    host = None
    if kind == "grid":
        import torch
        from isosurface_b200 import _lib
        lib = _lib.load()
        nl = min(size + 1, 257)
        host = make_field(lib, torch, 0, wl, 0, nl).cpu().numpy().reshape(nl, size, size)
    # W warm-ups + K steps of a bounded sample; keep the whole run within a few minutes
    per = max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(wl, host, target_s=per)
        if i >= args.warmup:
            vals.append(last)
    v = float(np.mean([x["value"] for x in vals]))
    vox_per_step = float(size) ** 3
    line = {"impl": "reference", "metric": "MarchingCubes Gvoxels/s", "value": v, "unit": "Gvoxels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": vox_per_step / (v * 1e9) * 1e3, "higher_is_better": True,
            "scaling": "weak" if (args.gpus == 1 or args.weak) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mtris_per_s": float(np.mean([x["mtris_per_s"] for x in vals])),
            "config": {"workload": wl, "size": size, "note": "ms_per_step extrapolated from the bounded sample to the full grid"},
            "cpu_baseline": {"value": v, "unit": "Gvoxels/s", "cores": 1, "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def run_ours(args):
    import torch
    import torch.distributed as dist
    import isosurface_b200 as iso
    from isosurface_b200 import _build, _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun --nproc-per-node %d" % (args.gpus, args.gpus))
    torch.cuda.set_device(local)
    if world > 1:
        # keep stdout to the one JSON line: a box-level NCCL_DEBUG=VERSION would print a banner there
        os.environ["NCCL_DEBUG"] = os.environ.get("ISOMC_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not _lib.LIB_PATH.exists():
        if rank == 0:
            _build.build_library()
        if world > 1:
            dist.barrier()
    lib = _lib.load()
    wl = args.workload
    size, kind, field, seed = WORKLOADS[wl]
    hbm_peak, peak_src = peaks()

    z0, z1 = slab_range(size, rank, world)
    ghost = 1 if z0 > 0 else 0
    n_layers = (z1 - z0) + ghost + 1
    h = C.c_void_p()
    _lib.check(lib.isomc_slab_create(size, z0, z1, local, C.byref(h)))
    stream = torch.cuda.Stream(device=local)
    _lib.check(lib.isomc_set_stream(h, C.c_void_p(stream.cuda_stream)), h)
    grid = None
    prog = None
    if kind == "grid":
        grid = make_field(lib, torch, local, wl, z0 - ghost, n_layers)
    else:
        sys.path.insert(0, str(ROOT / "tests"))
        from helpers import iso_source
        from isosurface_b200.source import encode_program
        prog = encode_program(iso_source(field))
    d_tot = C.c_void_p()
    _lib.check(lib.isomc_slab_totals_device(h, C.byref(d_tot)), h)
    mine = torch.as_tensor(CudaArray(d_tot.value, 3, "<i8"), device="cuda:%d" % local)
    gathered = torch.zeros(3 * world, dtype=torch.int64, device="cuda:%d" % local)

    def step():
        with torch.cuda.stream(stream):
            if world == 1:  # the plain single-GPU entry points (what MarchingCubes.extract calls)
                if kind == "grid":
                    _lib.check(lib.isomc_enqueue_grid_device(h, C.c_void_p(grid.data_ptr())), h)
                else:
                    _lib.check(lib.isomc_enqueue_sdf(h, prog.ctypes.data, len(prog)), h)
                _lib.check(lib.isomc_finish(h), h)
                return
            if kind == "grid":
                _lib.check(lib.isomc_slab_count_grid_device(h, C.c_void_p(grid.data_ptr())), h)
            else:
                _lib.check(lib.isomc_slab_count_sdf(h, prog.ctypes.data, len(prog)), h)
            dist.all_gather_into_tensor(gathered, mine)
            _lib.check(lib.isomc_slab_emit_gathered(h, C.c_void_p(gathered.data_ptr()), rank, world), h)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    st = _lib.Stats()
    _lib.check(lib.isomc_stats_get(h, C.byref(st)), h)
    launches_per_step = int(st.kernel_launches)

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record(stream)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
    ms_dev = ev0.elapsed_time(ev1)
    tmax = torch.tensor([ms_dev], dtype=torch.float64, device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_step = float(tmax.item()) / args.steps

    # per-kernel breakdown (CUDA events on the launching stream, recorded inside the library)
    _lib.check(lib.isomc_set_profiling(h, 1), h)
    prof = []
    for _ in range(min(args.steps, 10)):
        step()
        _lib.check(lib.isomc_stats_get(h, C.byref(st)), h)
        prof.append((st.ms_sign, st.ms_count, st.ms_scan, st.ms_emit, st.ms_total))
    _lib.check(lib.isomc_set_profiling(h, 0), h)
    prof = np.array(prof, dtype=np.float64).mean(axis=0)

    counts = torch.tensor([st.n_vertices, st.n_triangles, st.n_active_cells, st.n_samples], dtype=torch.int64,
                          device="cuda:%d" % local)
    if world > 1:
        dist.all_reduce(counts)
    V, T, A, S_all = [int(x) for x in counts.tolist()]
    S = size * size * (size + 1)
    voxels = float(size) ** 3
    b_alg = 4 * S + 12 * V + 12 * T
    t_s = ms_step * 1e-3
    value = voxels / t_s / 1e9
    line = {
        "metric": "MarchingCubes Gvoxels/s", "value": value, "unit": "Gvoxels/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if (world == 1 or args.weak) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mtris_per_s": T / t_s / 1e6, "gcells_per_s": float(size - 1) ** 2 * size / t_s / 1e9,
        "config": {"workload": wl, "size": size, "source": kind, "vertices": V, "triangles": T, "active_cells": A,
                   "active_fraction": A / (float(size - 1) ** 2 * size), "parallelism": "zslab%d" % world,
                   "l2": "input %.0f MB per GPU %s 126 MB L2; no explicit flush" % (4 * S / world / 1e6, ">" if 4 * S / world > 126e6 else "<"),
                   "timing": "CUDA events on the extraction stream, max over ranks; wall %.3f ms/step" % (wall / args.steps * 1e3)},
        "gpu_launches": launches_per_step * args.steps,
    }
    # roofline of the dominant kernel and of the whole extract
    # kernel names of the path in use: the tile path (default) or the older active-cell-list kernels (ISOMC_PATH=list)
    tile = os.environ.get("ISOMC_PATH", "tile") != "list"
    k_first, k_count, k_emit = ("k_tile_count", None, "k_tile_emit") if tile else ("k_sign", "k_count_list", "k_emit_list")
    # k_tile_count / k_sign read every sample once; the emit kernels write vertices and triangles
    kern = {k_first: (prof[0], 4 * S / world), "k_scan_rows": (prof[2], 0), k_emit: (prof[3], (12 * V + 12 * T) / world)}
    if k_count:
        kern[k_count] = (prof[1], 0)
    dom = max((k_first, k_emit), key=lambda k: kern[k][0])
    traffic = None
    tp = ROOT / "profiles" / "traffic.json"
    if tp.exists():
        try:
            traffic = json.loads(tp.read_text()).get(wl, {}).get(dom)
        except Exception:
            traffic = None
    if kind == "grid":
        ach = kern[dom][1] / (kern[dom][0] * 1e-3) / 1e9 if kern[dom][0] > 0 else 0.0
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kern[dom][1], "ms_per_launch": kern[dom][0]}
        ach_all = b_alg / world / t_s / 1e9
        line["roofline_extract"] = {"bound": "hbm", "scope": "whole extract (%s)" % "+".join(kern),
                                    "achieved": ach_all * world, "per_gpu": ach_all, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": ach_all / hbm_peak, "algorithmic_bytes": b_alg,
                                    "formula": "4*S + 12*V + 12*T"}
    else:
        ops = SDF_OPS.get(field, 20)
        peak = 148 * 128 * 1.965e9 / 1e12
        ach = S * ops / (prof[0] * 1e-3) / 1e12 if prof[0] > 0 else 0.0
        line["roofline"] = {"bound": "fp32", "kernel": k_first + "<Sdf>", "achieved": ach, "peak": peak, "unit": "Tlane-op/s",
                            "frac": ach / peak, "traffic": None, "peak_source": "148 SM x 128 lanes x 1.965 GHz, non-FMA",
                            "ops_per_sample": ops}
    line["kernels_ms"] = {k: v[0] for k, v in kern.items()}
    line["kernels_ms"]["sum"] = prof[4]
    line["config"]["path"] = "tile path (TMA-staged count, plane emission)" if tile else "active-cell list"
    line["clocks"] = clk.summary()

    # ---- e2e through the public API with HOST buffers (H2D of the grid + D2H of the mesh inside the timed region)
    if world == 1 and rank == 0 and not args.no_e2e and 4 * S > (8 << 30):
        line["e2e"] = None  # a > 8 GiB pinned host copy of the grid is not attempted
    elif world == 1 and rank == 0 and not args.no_e2e:
        mc = iso.MarchingCubes(size, device=local)
        if kind == "grid":
            hgrid = torch.empty(grid.numel(), dtype=torch.float32, pin_memory=True)
            hgrid.copy_(grid)
            hxyz = torch.empty(3 * V + 16, dtype=torch.float32, pin_memory=True)
            hidx = torch.empty(3 * T + 16, dtype=torch.int32, pin_memory=True)
            torch.cuda.synchronize()

            def e2e_step():
                _lib.check(lib.isomc_extract_grid_host_to(mc._h, C.c_void_p(hgrid.data_ptr()), C.c_void_p(hxyz.data_ptr()), V + 5,
                                                          C.c_void_p(hidx.data_ptr()), T + 5), mc._h)
            h2d = 4 * S
        else:
            src = iso.Sampler(iso_source(field))
            sink = iso.ArrayMesh()

            def e2e_step():
                mc.extract(src, sink)
            h2d = 16 * len(prog)
        n_e2e = max(1, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        dt = (time.perf_counter() - t0) / n_e2e
        line["e2e"] = {"value": voxels / dt / 1e9, "unit": "Gvoxels/s", "h2d_bytes_per_step": h2d,
                       "d2h_bytes_per_step": 12 * V + 12 * T, "ms_per_step": dt * 1e3, "steps": n_e2e,
                       "api": "isomc_extract_grid_host_to: host lattice -> host mesh, z-chunk pipelined (pinned host buffers)" if kind == "grid"
                       else "MarchingCubes.extract(Sampler(source), ArrayMesh())"}
        mc.close()
    elif world > 1:
        line["e2e"] = None

    # ---- CPU baseline beside the GPU number (rank 0, N=1 only)
    if world == 1 and rank == 0 and not args.no_cpu:
        host = None
        if kind == "grid":
            nl = min(size + 1, 257)
            host = grid[: nl * size * size].cpu().numpy().reshape(nl, size, size)
        cb = cpu_baseline(wl, host, target_s=args.cpu_seconds)
        line["cpu_baseline"] = cb
        line["config"]["host_cores"] = os.cpu_count()

    lib.isomc_destroy(h)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    args.weak = args.workload == "auto"
    if args.workload == "auto":  # 512^3 voxels per GPU of the same fBm family (weak scaling); fbm512 at N = 1
        args.workload = {1: "fbm512", 2: "fbm644", 4: "fbm812", 8: "fbm1024"}.get(args.gpus, "fbm1024")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
