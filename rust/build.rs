// build.rs for dusk-schnorr with the `cuda` feature (replaces /root/reference/build.rs, which only prints a
// deprecation warning).  NOT compiled in this image -- see rust/README.md.
use std::{env, path::PathBuf, process::Command};

fn main() {
    println!("cargo:rerun-if-changed=build.rs");
    if env::var("CARGO_FEATURE_CUDA").is_err() {
        return;
    }
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libschnorr_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", // B200 only: no multi-arch fatbin, no fallback path
            "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC", "-o",
        ])
        .arg(&lib)
        .arg("cuda/schnorr_b200.cu")
        .status()
        .expect("nvcc not found: the `cuda` feature needs the CUDA 12.9+ toolkit");
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=schnorr_b200");
    println!("cargo:rerun-if-changed=cuda");
}
