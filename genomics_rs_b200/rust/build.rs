// build.rs -- NOT COMPILED IN THIS ENVIRONMENT (no cargo/rustc in the image); shipped as the binding a
// maintainer of nlaha/genomics-rs would add.  Compiles the CUDA sources for sm_100a into a static library and
// links it, exactly the nvcc line genomics_rs_b200/build.py uses for the shared library.
use std::{env, path::PathBuf, process::Command};

fn main() {
    let out = PathBuf::from(env::var("OUT_DIR").unwrap());
    let csrc = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("gxalign/csrc");
    let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());
    let mut objs = Vec::new();
    // (source, object name, extra defines): the fill kernel is instantiated once per (K, recurrence form)
    let mut units: Vec<(&str, String, Vec<String>)> = vec![
        ("gx_api.cu", "gx_api.o".into(), vec![]),
        ("gx_k0.cu", "gx_k0.o".into(), vec![]),
    ];
    for k in [4, 8, 16] {
        for c in [0, 1] {
            units.push(("gx_fill_inst.cu", format!("gx_fill_k{k}_c{c}.o"), vec![format!("-DGX_INST_K={k}"), format!("-DGX_INST_CHAIN={c}")]));
        }
    }
    for (src, name, defs) in &units {
        let obj = out.join(name);
        let ok = Command::new(&nvcc)
            .args(["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
            .args(["-Xcompiler", "-fPIC"])
            .args(defs)
            .args(["-c", "-o"])
            .arg(&obj)
            .arg(csrc.join(src))
            .status()
            .expect("nvcc not found")
            .success();
        assert!(ok, "nvcc failed on {src}");
        objs.push(obj);
        println!("cargo:rerun-if-changed={}", csrc.join(src).display());
    }
    let lib = out.join("libgxalign.a");
    assert!(Command::new("ar").arg("crs").arg(&lib).args(&objs).status().unwrap().success());
    println!("cargo:rustc-link-search=native={}", out.display());
    println!("cargo:rustc-link-lib=static=gxalign");
    println!("cargo:rustc-link-search=native=/usr/local/cuda/lib64");
    println!("cargo:rustc-link-lib=cudart");
    println!("cargo:rustc-link-lib=stdc++");
}
