// build.rs -- compiles EVERY CUDA source under isosurface_b200/csrc for sm_100a with nvcc into libisomc_b200.so and links it.
// Not run in this repo's environment (no Rust toolchain); the source list and the flags are those of isosurface_b200/_build.py
// (tests/test_abi.py::test_rust_build_script_lists_every_source keeps the two in step).
use std::{env, fs, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("..");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = root.join("isosurface_b200").join("csrc");
    let lib = out.join("libisomc_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    // the translation units of the library (same list as _build.SOURCES)
    let sources = ["isomc_kernels.cu", "isomc_list_kernels.cu", "isomc_tile_kernels.cu", "isomc_points.cu", "isomc_api.cu",
                   "isomc_sharded.cu"];
    let mut cmd = Command::new(nvcc);
    cmd.args(&["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
               "-Xcompiler", "-fPIC", "-shared", "-ldl", "-o"]).arg(&lib);
    for s in &sources { cmd.arg(csrc.join(s)); }
    let status = cmd.status().expect("nvcc not found: the B200 path has no CPU fallback");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=isomc_b200");
    // every header and source the library is built from
    for entry in fs::read_dir(&csrc).unwrap() {
        let p = entry.unwrap().path();
        if matches!(p.extension().and_then(|e| e.to_str()), Some("cu") | Some("cuh") | Some("h")) {
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
    println!("cargo:rerun-if-changed={}", root.join("include").join("isomc.h").display());
}
