// Compiles the CUDA sources of libfhe_b200 with nvcc for sm_100a and links the result.
// FHE_B200_CSRC overrides the source directory (default: ../../fhe_study_b200/csrc).
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("FHE_B200_CSRC").unwrap_or_else(|_| "../../fhe_study_b200/csrc".into()));
    println!("cargo:rerun-if-changed={}", csrc.display());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let mut objs = Vec::new();
    for entry in std::fs::read_dir(&csrc).expect("csrc directory") {
        let p = entry.unwrap().path();
        if p.extension().map_or(false, |x| x == "cu") {
            let o = out.join(p.file_stem().unwrap()).with_extension("o");
            let ok = Command::new(&nvcc)
                .args(["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
                       "--fmad=false", "-Xcompiler", "-fPIC,-fvisibility=hidden", "-c"])
                .arg(&p).arg("-o").arg(&o)
                .status().expect("nvcc").success();
            assert!(ok, "nvcc failed on {}", p.display());
            objs.push(o);
        }
    }
    let lib = out.join("libfhe_b200.so");
    let ok = Command::new(&nvcc).arg("-shared").arg("-o").arg(&lib).args(&objs)
        .args(["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
        .status().expect("nvcc link").success();
    assert!(ok, "link failed");
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=dylib=fhe_b200");
}
