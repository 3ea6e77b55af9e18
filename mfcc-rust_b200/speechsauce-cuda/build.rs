//! Compiles the CUDA sources in ../csrc with `nvcc -arch=sm_100a` into libspeechsauce_cuda.so and
//! tells cargo to link it.  NOTE: there is no Rust toolchain in the image this repository is
//! developed in, so this file (like the rest of the crate) is written against the C ABI in
//! ../../include/speechsauce_cuda.h and has not been compiled; the same sources are built and
//! tested through ../Makefile and the ctypes mirror in ../speechsauce_b200.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let csrc = manifest.join("../csrc");
    let include = manifest.join("../../include");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libspeechsauce_cuda.so");

    if env::var("CARGO_FEATURE_PREBUILT").is_ok() {
        let dir = env::var("SPEECHSAUCE_CUDA_LIB_DIR").unwrap_or_else(|_| manifest.join("../lib").display().to_string());
        println!("cargo:rustc-link-search=native={}", dir);
    } else {
        let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
        let sources = ["ss_api.cu", "ss_tables.cpp", "ss_kernels_generic.cu", "ss_kernels_elem.cu", "ss_kernels_fast512c.cu", "ss_kernels_pspec512.cu", "ss_kernels_stft512.cu", "ss_kernels_stft400m.cu",
                       "ss_kernels_stft2048.cu", "ss_kernels_stft2048m.cu"];
        let mut cmd = Command::new(&nvcc);
        cmd.args(["-O3", "-std=c++17", "-arch=sm_100a", "-lineinfo", "--expt-relaxed-constexpr", "-shared",
                  "-Xcompiler", "-fPIC,-fvisibility=hidden,-ffp-contract=off"])  // host tables: no fused a * b + c (band edges sit on integers)
            .arg(format!("-I{}", include.display()))
            .arg(format!("-I{}", csrc.display()))
            .arg("-o")
            .arg(&lib);
        for s in sources.iter() {
            cmd.arg(csrc.join(s));
            println!("cargo:rerun-if-changed={}", csrc.join(s).display());
        }
        for h in ["ss_internal.h", "ss_host.h", "fft_ct.cuh"].iter() {
            println!("cargo:rerun-if-changed={}", csrc.join(h).display());
        }
        let status = cmd.status().expect("failed to run nvcc (set NVCC=/path/to/nvcc)");
        assert!(status.success(), "nvcc failed");
        println!("cargo:rustc-link-search=native={}", out.display());
    }
    println!("cargo:rustc-link-lib=dylib=speechsauce_cuda");
    println!("cargo:rerun-if-changed={}", include.join("speechsauce_cuda.h").display());
}
