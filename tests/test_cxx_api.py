"""The C++ host mirror (include/isosurface.hpp) compiles against the C ABI and behaves: on a CPU box it
fails loudly with ISOMC_ERR_CUDA; on the GPU box it runs the reference's bench case (benches/isosurface.rs:21-31)."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _build(tmp_path, isolib):
    exe = tmp_path / "cxx_api_check"
    libdir = ROOT / "isosurface_b200"
    subprocess.run(["g++", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cxx_api_check.cpp"), "-o", str(exe),
                    "-L", str(libdir), "-lisomc_b200", "-Wl,-rpath," + str(libdir)], check=True)
    return exe


def test_cxx_mirror_compiles_and_fails_loudly_without_gpu(tmp_path, isolib):
    import torch
    exe = _build(tmp_path, isolib)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    if not torch.cuda.is_available():
        assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_cxx_mirror_runs_the_reference_bench_case(tmp_path, isolib):
    exe = _build(tmp_path, isolib)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    # Torus(0.25,0.1) at the origin, N=128: V=2974 T=5718; then csgA @0.5 at N=128 appended
    v, t = [int(x.split("=")[1]) for x in r.stdout.split()]
    assert v > 2974 and t > 5718
