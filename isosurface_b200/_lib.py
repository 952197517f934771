"""ctypes binding of libisomc_b200.so (the C ABI in include/isomc.h).

The library is the product; this module only loads it.  There is no fallback of any kind: if the
shared object is missing or has no usable CUDA device, calls fail loudly.
"""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
import os
LIB_PATH = Path(os.environ["ISOMC_LIB"]) if os.environ.get("ISOMC_LIB") else PKG / "libisomc_b200.so"  # (ISOMC_LIB: experiment builds)

OK = 0
ERR_BAD_ARG, ERR_CUDA, ERR_OOM, ERR_INDEX_OVERFLOW, ERR_UNSUPPORTED_SOURCE, ERR_NO_RESULT, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
ERR_BUFFER_TOO_SMALL = -8
_ERR_NAMES = {-1: "BAD_ARG", -2: "CUDA", -3: "OOM", -4: "INDEX_OVERFLOW", -5: "UNSUPPORTED_SOURCE", -6: "NO_RESULT", -7: "NCCL",
              -8: "BUFFER_TOO_SMALL"}

SDF_SPHERE, SDF_TORUS, SDF_CYLINDER, SDF_PRISM = 1, 2, 3, 4
SDF_UNION, SDF_INTERSECTION, SDF_DIFFERENCE = 16, 17, 18
SDF_TRANSLATE_PUSH, SDF_TRANSLATE_POP = 32, 33
FIELD_FBM, FIELD_GYROID, FIELD_SPHERE_UNION = 1, 2, 3

NODE_DTYPE = np.dtype([("op", "<u4"), ("a", "<f4"), ("b", "<f4"), ("c", "<f4")])


class Stats(C.Structure):
    _fields_ = [("n_vertices", C.c_uint64), ("n_triangles", C.c_uint64), ("n_active_cells", C.c_uint64),
                ("n_samples", C.c_uint64), ("n_cells", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
                ("ms_sign", C.c_float), ("ms_count", C.c_float), ("ms_scan", C.c_float), ("ms_emit", C.c_float),
                ("ms_total", C.c_float), ("kernel_launches", C.c_uint32), ("emit_reruns", C.c_uint32)]


class IsomcError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("isomc error %s (%d): %s" % (_ERR_NAMES.get(code, "?"), code, message))
        self.code = code


# every symbol include/isomc.h declares: (name, restype, argtypes)
_P, _U32, _I32, _U64 = C.c_void_p, C.c_uint32, C.c_int32, C.c_uint64
SIGNATURES = {
    "isomc_create": (_I32, [_U32, _I32, C.POINTER(_P)]),
    "isomc_destroy": (_I32, [_P]),
    "isomc_last_error": (C.c_char_p, [_P]),
    "isomc_version": (C.c_char_p, []),
    "isomc_device_count": (_I32, []),
    "isomc_extract_sdf": (_I32, [_P, _P, _U32]),
    "isomc_extract_sdf_directed": (_I32, [_P, _P, _U32]),
    "isomc_extract_grid_device": (_I32, [_P, _P]),
    "isomc_extract_grid_host": (_I32, [_P, _P]),
    "isomc_extract_grid_host_to": (_I32, [_P, _P, _P, _U64, _P, _U64]),
    "isomc_batch_create": (_I32, [_U32, _U32, _I32, C.POINTER(_P)]),
    "isomc_extract_sdf_batch": (_I32, [_P, _P, _P, _U32]),
    "isomc_extract_sdf_batch_directed": (_I32, [_P, _P, _P, _U32]),
    "isomc_batch_offsets": (_I32, [_P, _P, _P]),
    "isomc_extract_grid_batch_device": (_I32, [_P, _P, _U32]),
    "isomc_extract_grid_batch_host": (_I32, [_P, _P, _U32]),
    "isomc_points_sdf": (_I32, [_P, _P, _U32]),
    "isomc_points_sdf_directed": (_I32, [_P, _P, _U32]),
    "isomc_points_grid_device": (_I32, [_P, _P]),
    "isomc_points_grid_host": (_I32, [_P, _P]),
    "isomc_counts": (_I32, [_P, C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64)]),
    "isomc_layer_counts": (_I32, [_P, _P]),
    "isomc_device_buffers": (_I32, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "isomc_copy_out": (_I32, [_P, _P, _P]),
    "isomc_copy_out_interleaved_normals": (_I32, [_P, _P, _U32, C.c_float, _P, _P]),
    "isomc_copy_out_interleaved_normals_at": (_I32, [_P, _P, _U32, C.c_float, _U32, _P, _P]),
    "isomc_stats_get": (_I32, [_P, C.POINTER(Stats)]),
    "isomc_get_stream": (_I32, [_P, C.POINTER(_P)]),
    "isomc_set_stream": (_I32, [_P, _P]),
    "isomc_enqueue_grid_device": (_I32, [_P, _P]),
    "isomc_enqueue_sdf": (_I32, [_P, _P, _U32]),
    "isomc_enqueue_sdf_directed": (_I32, [_P, _P, _U32]),
    "isomc_finish": (_I32, [_P]),
    "isomc_reserve": (_I32, [_P, _U64, _U64]),
    "isomc_set_profiling": (_I32, [_P, _I32]),
    "isomc_slab_create": (_I32, [_U32, _U32, _U32, _I32, C.POINTER(_P)]),
    "isomc_slab_count_grid_device": (_I32, [_P, _P]),
    "isomc_slab_count_sdf": (_I32, [_P, _P, _U32]),
    "isomc_slab_count_sdf_directed": (_I32, [_P, _P, _U32]),
    "isomc_slab_totals": (_I32, [_P, C.POINTER(_U64 * 3)]),
    "isomc_slab_totals_device": (_I32, [_P, C.POINTER(_P)]),
    "isomc_slab_emit": (_I32, [_P, _U64, _U64]),
    "isomc_slab_emit_gathered": (_I32, [_P, _P, _U32, _U32]),
    "isomc_slab_enqueue_emit_gathered": (_I32, [_P, _P, _U32, _U32]),
    "isomc_slab_mailbox": (_I32, [_P, C.POINTER(_P)]),
    "isomc_slab_mailbox_ipc": (_I32, [_P, _P]),
    "isomc_slab_connect": (_I32, [_P, _U32, _U32, _P]),
    "isomc_slab_connect_ipc": (_I32, [_P, _U32, _U32, _P]),
    "isomc_slab_emit_exchanged": (_I32, [_P]),
    "isomc_slab_enqueue_emit_exchanged": (_I32, [_P]),
    "isomc_slab_extract_grid_exchanged": (_I32, [_P, _P]),
    "isomc_slab_enqueue_extract_grid_exchanged": (_I32, [_P, _P]),
    "isomc_sharded_create": (_I32, [_U32, _U32, _P, C.POINTER(_P)]),
    "isomc_sharded_destroy": (_I32, [_P]),
    "isomc_sharded_last_error": (C.c_char_p, [_P]),
    "isomc_sharded_uses_nccl": (_I32, [_P]),
    "isomc_sharded_uses_peer_memory": (_I32, [_P]),
    "isomc_sharded_slab": (_I32, [_P, _U32, C.POINTER(_U32), C.POINTER(_U32), C.POINTER(_U32), C.POINTER(_U32)]),
    "isomc_sharded_handle": (_I32, [_P, _U32, C.POINTER(_P)]),
    "isomc_sharded_extract_grid": (_I32, [_P, _P]),
    "isomc_sharded_extract_sdf": (_I32, [_P, _P, _U32]),
    "isomc_sharded_extract_sdf_directed": (_I32, [_P, _P, _U32]),
    "isomc_sharded_counts": (_I32, [_P, C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64)]),
    "isomc_sharded_rank_counts": (_I32, [_P, _U32, C.POINTER(_U64), C.POINTER(_U64), C.POINTER(_U64)]),
    "isomc_sharded_copy_out": (_I32, [_P, _P, _P]),
    "isomc_debug_cube_indices": (_I32, [_P, _P]),
    "isomc_debug_sample_sdf": (_I32, [_I32, _P, _U32, _P, _U64, _P]),
    "isomc_synth_field": (_I32, [_I32, _I32, _U32, _U64, _U32, _U32, _P]),
}

_lib = None


def load():
    """Load the shared library (no compute happens here, so this also works on a CPU-only box)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise ImportError(
                "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  isosurface_b200 has no CPU fallback." % LIB_PATH)
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the ABI is incomplete
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def check(rc, handle=None):
    if rc != OK:
        msg = load().isomc_last_error(handle)
        raise IsomcError(rc, msg.decode() if msg else "")
    return rc
