// Builds librimphony_b200.so from ../rimphony_b200/csrc with nvcc for sm_100a and links it
// (replaces leung-bessel/build.rs:7-11 and gsl-sys/build.rs:11-32 of the reference: no C
// Bessel code, no GSL).  NVCC overrides the compiler path.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let src = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../rimphony_b200/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = vec![];
    for unit in std::fs::read_dir(&src).unwrap().filter_map(|e| e.ok()) {
        let p = unit.path();
        if p.extension().map_or(false, |e| e == "cu") {
            let o = out.join(p.file_stem().unwrap()).with_extension("o");
            let ok = Command::new(&nvcc)
                .args(["-O3", "-std=c++17", "-lineinfo", "--expt-relaxed-constexpr", "-gencode",
                       "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-c", "-o"])
                .arg(&o)
                .arg(&p)
                .status()
                .unwrap()
                .success();
            assert!(ok, "nvcc failed on {:?}", p);
            objs.push(o);
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
    let lib = out.join("librimphony_b200.so");
    assert!(Command::new(&nvcc)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o"])
        .arg(&lib)
        .args(&objs)
        .arg("-lcudart")
        .status()
        .unwrap()
        .success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=rimphony_b200");
    println!("cargo:rerun-if-changed=../include/rimphony_b200.h");
}
