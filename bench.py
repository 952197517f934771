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
    spheres2048  2048^3 dense f32 grid, union of 64 spheres (C5)
    torus256 / csga256 / csgb256   256^3 implicit SDF evaluated on device (C2)
    sphere32 / torus128            the reference's CPU-sized cases (C1a / C1b)

Besides the headline `value`, every line of a default run (any N) carries the north star's STRONG-scaling configs as
records measured in the same job: `strong_2048` (C5, spheres2048) and `strong_1024` (C4, gyroid1024), each with the
N-GPU step time, the single-GPU (unsharded, rank 0) step time of the same job, the speed-up and the HBM-roofline
fraction per GPU.  `--no-strong` skips them.

N > 1 is launched by torchrun (one rank per GPU); ranks own z-slabs, exchange 3 x u64 totals with one
NCCL all-gather on the extraction stream and write globally numbered indices directly.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU oracle (a C restatement of the
reference algorithm; the Rust reference cannot be built in this image) on a bounded sample of the same
workload -- the only place besides tests/ and smoke() where oracle/ is executed; it never loads the product library.
"""
import argparse
import ctypes as C
import hashlib
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


HOST_ONLY_SOURCES = ("isomc_api.cu", "isomc_sharded.cu")  # entry points and plumbing: no device code


def kernel_source_sha():
    """sha256 over the sources that hold device code: what an ncu capture (profiles/traffic.json) is valid for.
    The GPU box has no .git."""
    h = hashlib.sha256()
    for f in sorted((ROOT / "isosurface_b200" / "csrc").glob("*")):
        if f.suffix in (".cu", ".cuh", ".h") and f.name not in HOST_ONLY_SOURCES:
            h.update(f.name.encode())
            h.update(f.read_bytes())
    return h.hexdigest()[:16]


def dram_traffic(workload, kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu capture -- only if it was taken on these very sources"""
    tp = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(tp.read_text()).get(workload, {})
        if d.get("src_sha") != kernel_source_sha():
            return None
        return d.get("kernels", {}).get(kernel)
    except Exception:
        return None


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
    """times the oracle on a bounded sample of the workload"""
    from oracle import oracle as O
    sys.path.insert(0, str(ROOT / "tests"))
    size, kind, field, seed = WORKLOADS[wl]
    threading_note = "1 thread (the reference is strictly single-threaded: no rayon / threads anywhere in src/)"
    if kind == "sdf":
        from helpers import oracle_prog
        prog = oracle_prog(field)
        t0 = time.perf_counter()
        xyz, idx, act = O.extract_sdf(size, prog, O.FAITHFUL)
        dt = time.perf_counter() - t0
        vox = float(size) ** 3
        return {"value": vox / dt / 1e9, "unit": "Gvoxels/s", "cores": 1, "kind": "port",
                "sample": "full %s extract (%d^3), oracle faithful-cost mode, %s, %.2f s" % (wl, size, threading_note, dt),
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
            "sample": "first %d of %d cell layers of %s, oracle faithful-cost mode, %s, %.2f s; "
                      "C restatement of the reference, not cargo bench" % (z2, size, wl, threading_note, dt),
            "mtris_per_s": len(idx) / 3 / dt / 1e6}


def base_config(wl, world):
    size, kind, field, seed = WORKLOADS[wl]
    S = size * size * (size + 1)
    return {"workload": wl, "size": size, "source": kind, "parallelism": "zslab%d" % world,
            "l2": "input %.0f MB per GPU %s 126 MB L2; no explicit flush" % (4 * S / world / 1e6, ">" if 4 * S / world > 126e6 else "<")}


def run_reference(args):
    """the reference's CPU implementation of the path (the oracle port), timed on the box's host cores.  No product code:
    the field comes from the oracle library's own host generator."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    wl = args.workload
    size, kind, field, seed = WORKLOADS[wl]
    host = None
    if kind == "grid":
        nl = min(size + 1, 161 if size <= 512 else 41)
        host = O.synth_field(FIELD_KIND[field], size, seed, 0, nl)
    # W warm-ups + K steps of a bounded sample; keep the whole run within a few minutes
    per = max(1.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, last = [], None
    for i in range(args.warmup + args.steps):
        last = cpu_baseline(wl, host, target_s=per)
        if i >= args.warmup:
            vals.append(last)
    v = float(np.mean([x["value"] for x in vals]))
    vox_per_step = float(size) ** 3
    cfg = base_config(wl, args.gpus)
    cfg.update({"host_cores": os.cpu_count(), "path": "CPU oracle (C restatement of the reference), faithful-cost mode",
                "note": "each step is a bounded sample (first cell layers of the workload); ms_per_step is extrapolated to the full grid",
                "timing": "time.perf_counter around the oracle call"})
    # the same keys as our arm's config: mesh size of the FULL workload from the committed oracle record of the benchmark fields
    try:
        gold = json.loads((ROOT / "tests" / "golden" / "full_hashes.json").read_text()).get(wl)
    except Exception:  # noqa: BLE001
        gold = None
    if gold and gold.get("z_cells") == size:
        cfg.update({"vertices": gold["vertices"], "triangles": gold["triangles"], "active_cells": gold["active_cells"],
                    "active_fraction": gold["active_cells"] / (float(size - 1) ** 2 * size)})
    cfg["kernel_src_sha"] = None  # (no kernels on this arm)
    line = {"impl": "reference", "metric": "MarchingCubes Gvoxels/s", "value": v, "unit": "Gvoxels/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": vox_per_step / (v * 1e9) * 1e3, "higher_is_better": True,
            "scaling": "weak" if (args.gpus == 1 or args.weak) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "mtris_per_s": float(np.mean([x["mtris_per_s"] for x in vals])),
            "config": cfg,
            "cpu_baseline": {"value": v, "unit": "Gvoxels/s", "cores": 1, "kind": "port", "sample": last["sample"]},
            "e2e": {"value": v, "unit": "Gvoxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit_line(line)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

class Runner:
    """one workload on this rank's slab (the whole lattice when world == 1): field, handle, step()"""

    def __init__(self, env, wl, world=None, rank=None, z_range=None):
        import torch
        from isosurface_b200 import _lib
        self.env, self.wl = env, wl
        self.torch, self._lib, self.lib = torch, _lib, env["lib"]
        self.world = env["world"] if world is None else world
        self.rank = env["rank"] if rank is None else rank
        self.local = env["local"]
        self.size, self.kind, self.field, self.seed = WORKLOADS[wl]
        t0 = time.perf_counter()
        z0, z1 = slab_range(self.size, self.rank, self.world) if z_range is None else z_range
        self.z0, self.z1 = z0, z1
        ghost = 1 if z0 > 0 else 0
        self.n_layers = (z1 - z0) + ghost + 1
        self.h = C.c_void_p()
        _lib.check(self.lib.isomc_slab_create(self.size, z0, z1, self.local, C.byref(self.h)))
        self.create_s = time.perf_counter() - t0
        self.stream = env["stream"]
        _lib.check(self.lib.isomc_set_stream(self.h, C.c_void_p(self.stream.cuda_stream)), self.h)
        self.grid = self.prog = None
        if self.kind == "grid":
            self.grid = make_field(self.lib, torch, self.local, wl, z0 - ghost, self.n_layers)
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            from helpers import iso_source
            from isosurface_b200.source import encode_program
            self.prog = encode_program(iso_source(self.field))
        d_tot = C.c_void_p()
        _lib.check(self.lib.isomc_slab_totals_device(self.h, C.byref(d_tot)), self.h)
        dev = "cuda:%d" % self.local
        self.mine = torch.as_tensor(CudaArray(d_tot.value, 3, "<i8"), device=dev)
        self.gathered = torch.zeros(3 * self.world, dtype=torch.int64, device=dev)
        # the totals exchange: peer stores into the ranks' mailboxes over NVLink (default), or an NCCL all-gather per step
        self.exchange = "nccl" if os.environ.get("ISOMC_EXCHANGE", "peer") == "nccl" else "peer"
        if self.world > 1 and self.exchange == "peer":
            import torch.distributed as dist
            from isosurface_b200.sharded import connect_peers_ipc
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            try:
                connect_peers_ipc(self.lib, self.h, self.rank, self.world)
            except _lib.IsomcError as e:  # (CUDA IPC not available between these processes): every rank falls back together
                print("[bench] rank %d: peer mailboxes unavailable (%s): NCCL all-gather instead" % (self.rank, e), file=sys.stderr)
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self.exchange = "nccl"

    def step(self):
        lib, _lib, h = self.lib, self._lib, self.h
        with self.torch.cuda.stream(self.stream):
            if self.world == 1:  # the plain single-GPU entry points (what MarchingCubes.extract calls)
                if self.kind == "grid":
                    _lib.check(lib.isomc_enqueue_grid_device(h, C.c_void_p(self.grid.data_ptr())), h)
                else:
                    _lib.check(lib.isomc_enqueue_sdf(h, self.prog.ctypes.data, len(self.prog)), h)
                _lib.check(lib.isomc_finish(h), h)
                return
            import torch.distributed as dist
            if self.kind == "grid" and self.exchange == "peer":  # count + exchange + emit: one launch sequence, one call
                _lib.check(lib.isomc_slab_extract_grid_exchanged(h, C.c_void_p(self.grid.data_ptr())), h)
                return
            if self.kind == "grid":
                _lib.check(lib.isomc_slab_count_grid_device(h, C.c_void_p(self.grid.data_ptr())), h)
            else:
                _lib.check(lib.isomc_slab_count_sdf(h, self.prog.ctypes.data, len(self.prog)), h)
            if self.exchange == "peer":
                _lib.check(lib.isomc_slab_emit_exchanged(h), h)
            else:
                dist.all_gather_into_tensor(self.gathered, self.mine)
                _lib.check(lib.isomc_slab_emit_gathered(h, C.c_void_p(self.gathered.data_ptr()), self.rank, self.world), h)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps, warmup):
        """(ms per step: CUDA events on the extraction stream, max over ranks; wall ms per step; first-step wall ms)"""
        torch = self.torch
        t0 = time.perf_counter()
        self.step()
        torch.cuda.synchronize()
        first_ms = (time.perf_counter() - t0) * 1e3
        for _ in range(max(2, warmup - 1)):
            self.step()
        self.barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record(self.stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        ev1.record(self.stream)
        self.barrier()
        wall = time.perf_counter() - t0
        tmax = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda:%d" % self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        return float(tmax.item()) / steps, wall / steps * 1e3, first_ms

    def stats(self):
        st = self._lib.Stats()
        self._lib.check(self.lib.isomc_stats_get(self.h, C.byref(st)), self.h)
        return st

    def profile(self, n):
        """per-kernel CUDA-event times recorded inside the library (mean over n steps)"""
        self._lib.check(self.lib.isomc_set_profiling(self.h, 1), self.h)
        prof = []
        for _ in range(n):
            self.step()
            st = self.stats()
            prof.append((st.ms_sign, st.ms_count, st.ms_scan, st.ms_emit, st.ms_total))
        self._lib.check(self.lib.isomc_set_profiling(self.h, 0), self.h)
        return np.array(prof, dtype=np.float64).mean(axis=0)

    def global_counts(self):
        st = self.stats()
        counts = self.torch.tensor([st.n_vertices, st.n_triangles, st.n_active_cells], dtype=self.torch.int64, device="cuda:%d" % self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(counts)
        return [int(x) for x in counts.tolist()]

    def close(self):
        self.lib.isomc_destroy(self.h)
        self.grid = None
        self.torch.cuda.empty_cache()


def strong_record(env, wl, steps, hbm_peak):
    """the north star's strong-scaling config `wl`: N-GPU step time, the single-GPU time of the same job, the speed-up.
    N > 1: timed twice -- z-slabs of equal thickness, then z-slabs of equal WORK cut from the per-layer active-cell counts of the
    first run (isomc_layer_counts -> sharded.balanced_slabs); `ms_per_step` is the better of the two, both are reported."""
    size = WORKLOADS[wl][0]
    S = size * size * (size + 1)
    r = Runner(env, wl)
    ms, wall_ms, _ = r.timed(steps, 3)
    V, T, A = r.global_counts()
    extra = {}
    if env["world"] > 1:
        import torch
        import torch.distributed as dist
        from isosurface_b200.sharded import balanced_slabs
        dev = "cuda:%d" % env["local"]
        mine = np.zeros(3 * (r.z1 - r.z0), np.uint64)
        r._lib.check(r.lib.isomc_layer_counts(r.h, mine.ctypes.data), r.h)
        layers = torch.zeros(size, dtype=torch.int64, device=dev)
        layers[r.z0:r.z1] = torch.from_numpy(mine.reshape(-1, 3)[:, 2].astype(np.int64)).to(dev)
        dist.all_reduce(layers)
        slabs = balanced_slabs(size, layers.cpu().numpy(), env["world"])
        r.close()
        extra = {"equal_split_ms_per_step": ms, "equal_split_slabs": [list(slab_range(size, k, env["world"])) for k in range(env["world"])]}
        if slabs != [slab_range(size, k, env["world"]) for k in range(env["world"])]:
            r = Runner(env, wl, z_range=slabs[env["rank"]])
            ms_b, _, _ = r.timed(steps, 3)
            Vb, Tb, Ab = r.global_counts()
            assert (Vb, Tb, Ab) == (V, T, A), "the balanced split must give the same mesh"
            extra.update({"balanced_split_ms_per_step": ms_b, "balanced_split_slabs": [list(s) for s in slabs],
                          "split": "balanced" if ms_b < ms else "equal"})
            ms = min(ms, ms_b)
    r.close()
    b_alg = 4 * S + 12 * V + 12 * T
    rec = {"workload": wl, "size": size, "n_gpus": env["world"], "steps": steps, "ms_per_step": ms,
           "gvoxels_per_s": float(size) ** 3 / (ms * 1e-3) / 1e9, "mtris_per_s": T / (ms * 1e-3) / 1e6,
           "vertices": V, "triangles": T, "active_cells": A,
           "roofline_extract": {"frac": b_alg / env["world"] / (ms * 1e-3) / 1e9 / hbm_peak, "per_gpu_gbs": b_alg / env["world"] / (ms * 1e-3) / 1e9,
                                "algorithmic_bytes": b_alg, "peak": hbm_peak},
           "timing": "CUDA events on the extraction stream, max over ranks"}
    rec.update(extra)
    if env["world"] == 1:
        rec["n1_ms_per_step"], rec["speedup"] = ms, 1.0
    else:
        import torch
        import torch.distributed as dist
        n1 = torch.zeros(1, dtype=torch.float64, device="cuda:%d" % env["local"])
        if env["rank"] == 0:  # the unsharded extract on one GPU of the same box, same job (the other ranks idle)
            r1 = Runner(env, wl, world=1, rank=0)
            n1[0] = r1.timed(steps, 3)[0]
            r1.close()
        dist.broadcast(n1, 0)
        rec["n1_ms_per_step"] = float(n1.item())
        rec["speedup"] = rec["n1_ms_per_step"] / ms
    return rec


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
        # stdout carries exactly one JSON line: whatever NCCL_DEBUG level the launcher asks for goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
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
    env = {"lib": lib, "world": world, "rank": rank, "local": local, "stream": torch.cuda.Stream(device=local)}

    t_cold = time.perf_counter()
    R = Runner(env, wl)
    with ClockSampler(local) as clk:
        ms_step, wall_ms, first_ms = R.timed(args.steps, args.warmup)
    cold_ms = R.create_s * 1e3 + first_ms
    launches_per_step = int(R.stats().kernel_launches)
    prof = R.profile(min(args.steps, 10))
    V, T, A = R.global_counts()
    S = size * size * (size + 1)
    voxels = float(size) ** 3
    b_alg = 4 * S + 12 * V + 12 * T
    t_s = ms_step * 1e-3
    value = voxels / t_s / 1e9
    cfg = base_config(wl, world)
    if world > 1:
        cfg["exchange"] = ("slab totals as peer stores into the ranks' mailboxes over NVLink (CUDA IPC), offset derived on the device"
                           if R.exchange == "peer" else "NCCL all-gather of 3 x u64 per rank on the extraction stream")
    cfg.update({"vertices": V, "triangles": T, "active_cells": A, "active_fraction": A / (float(size - 1) ** 2 * size),
                "timing": "CUDA events on the extraction stream, max over ranks; wall %.3f ms/step" % wall_ms})
    line = {
        "metric": "MarchingCubes Gvoxels/s", "value": value, "unit": "Gvoxels/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak" if (world == 1 or args.weak) else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "mtris_per_s": T / t_s / 1e6, "gcells_per_s": float(size - 1) ** 2 * size / t_s / 1e9,
        "config": cfg,
        "gpu_launches": launches_per_step * args.steps,
        # handle creation (all scratch allocations) + the first extract of the handle (count, size read-back, output
        # allocation, emission): what the reference's own bench times per iteration (benches/isosurface.rs:21-31)
        "cold_ms": cold_ms,
    }
    # roofline of the dominant kernel and of the whole extract
    # kernel names of the path in use: the tile path (default) or the older active-cell-list kernels (ISOMC_PATH=list)
    tile = os.environ.get("ISOMC_PATH", "list") == "tile"
    k_first, k_count, k_emit = ("k_tile_count", None, "k_tile_emit") if tile else ("k_sign", "k_count_list", "k_emit_list")
    # k_tile_count / k_sign read every sample once; the emit kernels write vertices and triangles
    kern = {k_first: (prof[0], 4 * S / world), "k_scan_rows": (prof[2], 0), k_emit: (prof[3], (12 * V + 12 * T) / world)}
    if k_count:
        kern[k_count] = (prof[1], 0)
    dom = max((k_first, k_emit), key=lambda k: kern[k][0])
    if kind == "grid":
        ach = kern[dom][1] / (kern[dom][0] * 1e-3) / 1e9 if kern[dom][0] > 0 else 0.0
        line["roofline"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                            "frac": ach / hbm_peak, "traffic": dram_traffic(wl, dom), "peak_source": peak_src,
                            "algorithmic_bytes_per_launch": kern[dom][1], "ms_per_launch": kern[dom][0]}
        ach_all = b_alg / world / t_s / 1e9
        line["roofline_extract"] = {"bound": "hbm", "scope": "whole extract (%s)" % "+".join(kern),
                                    "achieved": ach_all * world, "per_gpu": ach_all, "peak": hbm_peak, "unit": "GB/s",
                                    "frac": ach_all / hbm_peak, "algorithmic_bytes": b_alg,
                                    "formula": "4*S + 12*V + 12*T"}
        # the streaming stage in the stream (CUDA events inside the library): bytes of samples it reads per second
        line["sample_stream_gbs"] = 4 * S / world / (prof[0] * 1e-3) / 1e9 if prof[0] > 0 else None
    else:
        ops = SDF_OPS.get(field, 20)
        peak = 148 * 128 * 1.965e9 / 1e12
        ach = S * ops / (prof[0] * 1e-3) / 1e12 if prof[0] > 0 else 0.0
        line["roofline"] = {"bound": "fp32", "kernel": k_first + "<Sdf>", "achieved": ach, "peak": peak, "unit": "Tlane-op/s",
                            "frac": ach / peak, "traffic": None, "peak_source": "148 SM x 128 lanes x 1.965 GHz, non-FMA",
                            "ops_per_sample": ops}
    line["kernels_ms"] = {k: v[0] for k, v in kern.items()}
    line["kernels_ms"]["sum"] = prof[4]
    line["config"]["path"] = "tile path (TMA-staged count, plane emission; ISOMC_PATH=tile)" if tile else "active-cell list"
    line["config"]["kernel_src_sha"] = kernel_source_sha()
    line["clocks"] = clk.summary()

    # ---- e2e through the public API with HOST buffers (H2D of the grid + D2H of the mesh inside the timed region)
    if args.no_e2e or (kind == "grid" and 4 * S / world > (8 << 30)):
        line["e2e"] = None  # a > 8 GiB pinned host copy of the grid is not attempted
    elif world == 1:
        mc = iso.MarchingCubes(size, device=local)
        if kind == "grid":
            hgrid = torch.empty(R.grid.numel(), dtype=torch.float32, pin_memory=True)
            hgrid.copy_(R.grid)
            hxyz = torch.empty(3 * V + 16, dtype=torch.float32, pin_memory=True)
            hidx = torch.empty(3 * T + 16, dtype=torch.int32, pin_memory=True)
            torch.cuda.synchronize()

            def e2e_step():
                _lib.check(lib.isomc_extract_grid_host_to(mc._h, C.c_void_p(hgrid.data_ptr()), C.c_void_p(hxyz.data_ptr()), V + 5,
                                                          C.c_void_p(hidx.data_ptr()), T + 5), mc._h)
            h2d = 4 * S
        else:
            sys.path.insert(0, str(ROOT / "tests"))
            from helpers import iso_source
            src = iso.Sampler(iso_source(field))
            sink = iso.ArrayMesh()

            def e2e_step():
                mc.extract(src, sink)
            h2d = 16 * len(R.prog)
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
        if kind == "grid":
            # what the link alone takes for the same pinned lattice (no kernels, no copy-out): the floor of ms_per_step
            best = None
            for _ in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                R.grid.view(-1).copy_(hgrid, non_blocking=True)
                torch.cuda.synchronize()
                d = time.perf_counter() - t0
                best = d if best is None or d < best else best
            line["e2e"]["h2d_only_ms"] = best * 1e3
        mc.close()
    elif kind == "grid":
        # every rank: pinned host slab -> device, count, all-gather, emit, its part of the mesh -> pinned host buffers
        st = R.stats()
        hslab = torch.empty(R.grid.numel(), dtype=torch.float32, pin_memory=True)
        hslab.copy_(R.grid)
        dslab = torch.empty_like(R.grid)
        hxyz = np.empty(3 * int(st.n_vertices) + 16, np.float32)
        hidx = np.empty(3 * int(st.n_triangles) + 16, np.uint32)
        torch.cuda.cudart().cudaHostRegister(hxyz.ctypes.data, hxyz.nbytes, 0)
        torch.cuda.cudart().cudaHostRegister(hidx.ctypes.data, hidx.nbytes, 0)
        saved = R.grid
        R.grid = dslab

        def e2e_step():
            with torch.cuda.stream(R.stream):
                dslab.copy_(hslab, non_blocking=True)
            R.step()
            _lib.check(lib.isomc_copy_out(R.h, hxyz.ctypes.data, hidx.ctypes.data), R.h)
        n_e2e = max(1, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        R.barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda:%d" % local)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt = float(dt.item())
        R.grid = saved
        torch.cuda.cudart().cudaHostUnregister(hxyz.ctypes.data)
        torch.cuda.cudart().cudaHostUnregister(hidx.ctypes.data)
        line["e2e"] = {"value": voxels / dt / 1e9, "unit": "Gvoxels/s", "h2d_bytes_per_step": 4 * S + 4 * size * size * 2 * (world - 1),
                       "d2h_bytes_per_step": 12 * V + 12 * T, "ms_per_step": dt * 1e3, "steps": n_e2e,
                       "api": "per rank: pinned host slab -> device, isomc_slab_count / all-gather / isomc_slab_emit_gathered, "
                              "isomc_copy_out to pinned host buffers; wall clock, max over ranks"}
    else:
        line["e2e"] = None

    # ---- CPU baseline beside the GPU number (rank 0, N=1 only)
    if world == 1 and rank == 0 and not args.no_cpu:
        host = None
        if kind == "grid":
            nl = min(size + 1, 257)
            host = R.grid[: nl * size * size].cpu().numpy().reshape(nl, size, size)
        cb = cpu_baseline(wl, host, target_s=args.cpu_seconds)
        line["cpu_baseline"] = cb
        line["config"]["host_cores"] = os.cpu_count()
    R.close()

    # ---- the north star's strong-scaling configs, measured in the same job
    if not args.no_strong and args.weak:
        for key, swl in (("strong_2048", "spheres2048"), ("strong_1024", "gyroid1024")):
            try:
                line[key] = strong_record(env, swl, max(3, min(args.steps, 8)), hbm_peak)
            except Exception as e:  # (out of memory on a smaller part, ...): say so instead of losing the headline line
                line[key] = {"workload": swl, "error": repr(e)[:300]}
                if world > 1:
                    raise

    if rank == 0:
        emit_line(line)
    if world > 1:
        dist.destroy_process_group()


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: file descriptor 1 is pointed at stderr for the rest of the process, so that what
    libraries print there (NCCL writes its version banner to stdout at NCCL_DEBUG=VERSION whatever NCCL_DEBUG_FILE says) cannot
    get in front of it; emit_line() writes to the real stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)


def emit_line(line):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(WORKLOADS))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong_2048 / strong_1024 records of a default run")
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
