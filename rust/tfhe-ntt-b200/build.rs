// Compiles the CUDA sources with nvcc for sm_100a and links them (static cudart), as the
// north-star asks: "a thin C-ABI FFI layer compiled by build.rs with nvcc".  Mirrors what
// tfhe-rs-main_modified_b200/build.py does for the Python-hosted tests.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("tfhe-rs-main_modified_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let sources = [
        "ntt_engine.cu", "ntt_fast_solinas.cu", "ntt_fast_shoup64.cu", "ntt_fast_shoup32.cu",
        "ntt_fast_exact.cu", "ntt_pbs_solinas.cu", "capi_prime.cu", "capi_native.cu", "capi_product.cu",
        "capi_pbs.cu", "capi_custum_radix.cu",
    ];
    let mut objects = Vec::new();
    for src in sources {
        let obj = out.join(src.replace(".cu", ".o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-c"])
            .arg(csrc.join(src)).arg("-o").arg(&obj)
            .status().expect("nvcc not found: set NVCC");
        assert!(status.success(), "nvcc failed on {src}");
        objects.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libtfhe_ntt_b200.a");
    let status = Command::new("ar").arg("crs").arg(&lib).args(&objects).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=tfhe_ntt_b200");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=static=cudart_static");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=dl");
    println!("cargo:rustc-link-lib=dylib=rt");
    println!("cargo:rerun-if-changed={}", root.join("include/tfhe_ntt_b200.h").display());
}
