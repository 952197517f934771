// build.rs -- compiles the CUDA sources for sm_100a with nvcc and links them (plus cudart).
// Not run in this repo's environment (no Rust toolchain); mirrors isosurface_b200/_build.py.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("isosurface_b200").join("csrc");
    let lib = out.join("libisomc_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(nvcc)
        .args(&["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
                "-Xcompiler", "-fPIC", "-shared", "-o"])
        .arg(&lib)
        .arg(csrc.join("isomc_kernels.cu"))
        .arg(csrc.join("isomc_api.cu"))
        .status()
        .expect("nvcc not found: the B200 path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=isomc_b200");
    for f in &["isomc_kernels.cu", "isomc_api.cu", "isomc_device.cuh", "isomc_tables.h", "isomc_kernels.h"] {
        println!("cargo:rerun-if-changed={}", csrc.join(f).display());
    }
}
