// build.rs — compiles the CUDA library with nvcc for sm_100a and links it (north_star: "Host code stays
// Rust and calls CUDA through a thin extern "C" FFI (build.rs invoking nvcc -arch=sm_100a)").
// NOT compiled in this repo's CI: the build image has no Rust toolchain. Paths are relative to this crate.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("eth_tx_proof_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    // device headers embedded as source text for the NVRTC-compiled constraint programs (etp_jit.cu): same rule as
    // eth_tx_proof_b200/csrc/Makefile (jit_headers.inc = gl.cuh, powtable.cuh, quotient_rt.cuh as C string literals)
    let mut inc = String::new();
    for h in ["gl.cuh", "powtable.cuh", "quotient_rt.cuh"] {
        for line in std::fs::read_to_string(csrc.join(h)).unwrap().lines() {
            inc.push_str(&format!("\"{}\\n\"\n", line.replace('\\', "\\\\").replace('"', "\\\"")));
        }
        inc.push_str(",\n");
    }
    std::fs::write(out.join("jit_headers.inc"), inc).unwrap();
    let mut objs = vec![];
    for unit in ["etp_core", "etp_stark", "etp_shard", "etp_jit"] {
        let obj = out.join(format!("{unit}.o"));
        let status = Command::new(&nvcc)
            .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                   "-Xcompiler", "-fPIC", "-I"])
            .arg(&out)
            .args(["-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(format!("{unit}.cu")))
            .status()
            .expect("nvcc not found");
        assert!(status.success(), "nvcc failed on {unit}.cu");
        objs.push(obj);
    }
    let lib = out.join("libetp_b200.a");
    let status = Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=etp_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=nvrtc");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.display());
    println!("cargo:rerun-if-changed={}", root.join("include/etp_b200.h").display());
}
