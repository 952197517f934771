"""The C-ABI library loads on a CPU-only box, exports every symbol include/isomc.h declares, and fails
loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "isomc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(isomc_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(isolib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(isolib, s), "libisomc_b200.so does not export %s" % s


def test_python_binding_covers_the_header(isolib):
    from isosurface_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_version_string(isolib):
    assert b"sm_100a" in isolib.isomc_version()


def test_no_cpu_fallback_without_device(isolib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from isosurface_b200 import _lib
    h = C.c_void_p()
    rc = isolib.isomc_create(32, 0, C.byref(h))
    assert rc == _lib.ERR_CUDA and not h.value
    assert b"no CPU fallback" in isolib.isomc_last_error(None)
    import isosurface_b200 as iso
    with pytest.raises(_lib.IsomcError) as ei:
        iso.MarchingCubes(32)
    assert ei.value.code == _lib.ERR_CUDA


def test_bad_arguments_are_reported_not_crashed(isolib):
    from isosurface_b200 import _lib
    h = C.c_void_p()
    assert isolib.isomc_create(0, 0, C.byref(h)) == _lib.ERR_BAD_ARG      # reference underflows at size 0
    assert isolib.isomc_create(100000, 0, C.byref(h)) == _lib.ERR_BAD_ARG
    assert isolib.isomc_slab_create(64, 10, 10, 0, C.byref(h)) == _lib.ERR_BAD_ARG
    assert isolib.isomc_slab_create(64, 0, 65, 0, C.byref(h)) == _lib.ERR_BAD_ARG
    assert isolib.isomc_counts(None, None, None, None) == _lib.ERR_BAD_ARG
    assert isolib.isomc_destroy(None) == _lib.OK


def test_product_never_touches_the_oracle():
    """the product path must not import/link/execute anything under oracle/"""
    for p in list((ROOT / "isosurface_b200").rglob("*.py")) + list((ROOT / "isosurface_b200" / "csrc").glob("*")) + \
            list((ROOT / "include").glob("*")):
        if p.is_file() and p.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp"):
            text = p.read_text()
            assert "oracle" not in text.lower() or p.name == "isomc.h" and False, "%s mentions the oracle" % p


def test_rust_build_script_lists_every_source():
    """rust/build.rs is never compiled here (no cargo in the image): at least its source list must be the library's"""
    from isosurface_b200 import _build
    text = (ROOT / "rust" / "build.rs").read_text()
    for src in _build.SOURCES:
        assert '"%s"' % src in text, "rust/build.rs does not compile %s" % src
    assert set(re.findall(r'"(isomc_[a-z_]+\.cu)"', text)) == set(_build.SOURCES)
    for flag in ("arch=compute_100a,code=sm_100a", "-fmad=false"):
        assert flag in text


def test_rust_and_cxx_mirrors_bind_only_declared_symbols():
    """every isomc_* function the Rust shim declares / the C++ mirror calls exists in include/isomc.h; the sharded, slab and
    batch entry points are bound in both"""
    syms = set(declared_symbols())
    rust = (ROOT / "rust" / "src" / "lib.rs").read_text()
    rust_fns = set(re.findall(r"\bfn (isomc_[a-z_0-9]+)\s*\(", rust))
    assert rust_fns and rust_fns <= syms, sorted(rust_fns - syms)
    cxx = (ROOT / "include" / "isosurface.hpp").read_text()
    cxx_fns = set(re.findall(r"\b(isomc_[a-z_0-9]+)\s*\(", cxx))
    assert cxx_fns <= syms, sorted(cxx_fns - syms)
    for need in ("isomc_sharded_create", "isomc_sharded_extract_grid", "isomc_sharded_counts", "isomc_sharded_copy_out",
                 "isomc_extract_sdf_batch", "isomc_extract_sdf_directed"):
        assert need in rust_fns and need in cxx_fns, need
    assert "isomc_slab_emit_exchanged" in rust_fns and "extract_directed" not in rust  # Directed is the D instantiation of extract
