// build.rs -- added to the rulinalg crate root by the integration.
// Compiles the B200 kernels with nvcc for sm_100a into a static library and links it together with
// the CUDA runtime.  NOT exercised in this repository's image (no rustc/cargo); the same sources are
// built into librla_b200.so by rulinalg_b200/csrc/Makefile for the parity tests.
use std::env;
use std::path::PathBuf;
use std::process::Command;

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("RLA_B200_CSRC").unwrap_or_else(|_| "rla_b200/csrc".into()));
    let cuda = env::var("CUDA_HOME").unwrap_or_else(|_| "/usr/local/cuda".into());
    // every source rulinalg_b200/csrc/Makefile builds (tests/test_abi.py keeps the two lists identical)
    let srcs = ["api.cu", "host.cu", "multi.cu", "dgemm.cu", "sgemm.cu", "lu.cu", "solve.cu", "fill.cu", "gemv.cu",
                "cholesky.cu", "peak.cu"];
    let mut objs = Vec::new();
    for s in srcs.iter() {
        let obj = out.join(format!("{}.o", s));
        let st = Command::new(format!("{}/bin/nvcc", cuda))
            .args(&["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                    "-Xcompiler", "-fPIC", "-c"])
            .arg(csrc.join(s))
            .arg("-o")
            .arg(&obj)
            .status()
            .expect("nvcc not found");
        assert!(st.success(), "nvcc failed on {}", s);
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(s).display());
    }
    let lib = out.join("librla_b200.a");
    let st = Command::new("ar").arg("crs").arg(&lib).args(&objs).status().expect("ar");
    assert!(st.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-search=native={}/lib64", cuda);
    println!("cargo:rustc-link-lib=static=rla_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rustc-link-lib=dylib=pthread");   // staging pool / drainer / per-GPU issue threads (host.cu, multi.cu)
    // no NCCL: inside one process the multi-GPU exchange is peer memory (multi.cu); NCCL is only used by the
    // one-process-per-GPU Python driver (rulinalg_b200/sharded*.py)
    for h in ["common.cuh", "context.cuh"].iter() {
        println!("cargo:rerun-if-changed={}", csrc.join(h).display());
    }
}
