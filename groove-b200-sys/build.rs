// Builds libgroove_b200.so with nvcc for sm_100a and links it (north_star: "nvcc invoked from build.rs").
// There is no CPU fallback: without nvcc the build fails.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let manifest = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap());
    let repo = manifest.parent().expect("groove-b200-sys sits inside the groove-b200 repository").to_path_buf();
    let csrc = repo.join("groove_b200").join("csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let lib = out.join("libgroove_b200.so");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".to_string());
    let status = Command::new(&nvcc)
        .args([
            "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--shared",
            "-Xcompiler", "-fPIC,-O2", "--cudart", "shared",
        ])
        .arg("-I").arg(&csrc)
        .arg("-o").arg(&lib)
        .arg(csrc.join("engine.cu"))
        .status()
        .unwrap_or_else(|e| panic!("could not run {nvcc}: {e} (groove-b200 has no CPU fallback)"));
    assert!(status.success(), "nvcc failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=groove_b200");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", repo.join("include").join("groove_b200.h").display());
    println!("cargo:rerun-if-env-changed=NVCC");
}
