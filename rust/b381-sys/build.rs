// Build libb381 for sm_100a with nvcc (via the `cc` crate's CUDA mode), or link a prebuilt one.
use std::env;
use std::path::PathBuf;

fn main() {
    if let Ok(dir) = env::var("B381_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=b381");
        return;
    }
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("plonky2-bls12-381-pairing_b200/csrc");
    println!("cargo:rerun-if-changed={}", csrc.display());
    cc::Build::new()
        .cuda(true)
        .cudart("shared")
        .flag("-gencode").flag("arch=compute_100a,code=sm_100a")
        .flag("-O3").flag("-lineinfo").flag("-std=c++17")
        .include(root.join("include"))
        .file(csrc.join("kernels.cu"))
        .compile("b381");
    println!("cargo:rustc-link-lib=dylib=cudart");
}
