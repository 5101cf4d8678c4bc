// build.rs — compiles the CUDA library with nvcc for sm_100a and links it (north_star: "Host code stays
// Rust and calls CUDA through a thin extern "C" FFI (build.rs invoking nvcc -arch=sm_100a)").
// NOT compiled in this repo's CI: the build image has no Rust toolchain. Paths are relative to this crate.
// The translation units come from eth_tx_proof_b200/csrc/units.txt, the same list the Makefile builds: .cu files go
// through nvcc, .cpp files (the host-side Poseidon of the Fiat-Shamir transcript) through the host C++ compiler.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
    let csrc = root.join("eth_tx_proof_b200/csrc");
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "nvcc".into());
    let cxx = env::var("CXX").unwrap_or_else(|_| "g++".into());
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
    let units = std::fs::read_to_string(csrc.join("units.txt")).expect("eth_tx_proof_b200/csrc/units.txt");
    let mut objs = vec![];
    for unit in units.split_whitespace() {
        let src = csrc.join(unit);
        let stem = unit.rsplit_once('.').map(|(s, _)| s).unwrap_or(unit);
        let obj = out.join(format!("{stem}.o"));
        let status = if unit.ends_with(".cu") {
            Command::new(&nvcc)
                .args(["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "--use_fast_math",
                       "-Xcompiler", "-fPIC", "-I"])
                .arg(&out)
                .args(["-c", "-o"])
                .arg(&obj)
                .arg(&src)
                .status()
                .expect("nvcc not found")
        } else {
            Command::new(&cxx)
                .args(["-O3", "-std=c++17", "-fPIC", "-c", "-o"])
                .arg(&obj)
                .arg(&src)
                .status()
                .expect("host C++ compiler not found")
        };
        assert!(status.success(), "compiling {unit} failed");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", src.display());
    }
    let lib = out.join("libetp_b200.a");
    let _ = std::fs::remove_file(&lib);
    let status = Command::new("ar").arg("rcs").arg(&lib).args(&objs).status().unwrap();
    assert!(status.success());
    if let Ok(cuda) = env::var("CUDA_HOME") {
        println!("cargo:rustc-link-search=native={cuda}/lib64");
    } else {
        println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    }
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=etp_b200");
    println!("cargo:rustc-link-lib=dylib=cudart");
    println!("cargo:rustc-link-lib=dylib=nvrtc");
    println!("cargo:rustc-link-lib=dylib=stdc++");
    println!("cargo:rerun-if-changed={}", csrc.join("units.txt").display());
    println!("cargo:rerun-if-changed={}", root.join("include/etp_b200.h").display());
    for h in std::fs::read_dir(&csrc).unwrap().flatten() {
        let p = h.path();
        if p.extension().map_or(false, |e| e == "cuh" || e == "h") {
            println!("cargo:rerun-if-changed={}", p.display());
        }
    }
}
