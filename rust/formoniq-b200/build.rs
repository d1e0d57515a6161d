// build.rs — compiles the CUDA sources with nvcc for sm_100a and links the
// result, exactly what `python -m formoniq_b200.build` does in this repo.
use std::{env, path::PathBuf, process::Command};

fn main() {
  let root = PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../..");
  let csrc = root.join("formoniq_b200/csrc");
  let out = PathBuf::from(env::var("OUT_DIR").unwrap());
  let nvcc = env::var("NVCC").unwrap_or_else(|_| "/usr/local/cuda/bin/nvcc".into());

  // 1. element-tape generator (host C++) -> elmat_gen.cuh
  let gen = out.join("gen_elmat");
  assert!(Command::new("g++").args(["-O1", "-std=c++17"]).arg(csrc.join("gen_elmat.cpp")).arg("-o").arg(&gen)
    .status().unwrap().success());
  let generated = Command::new(&gen).output().unwrap();
  std::fs::write(csrc.join("elmat_gen.cuh"), generated.stdout).unwrap();

  // 2. nvcc: one shared library, static CUDA runtime, no CPU fallback
  let lib = out.join("libformoniq_b200.so");
  let mut cmd = Command::new(nvcc);
  cmd.args(["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
            "-Xcompiler", "-fPIC", "-shared", "-cudart", "static", "-o"]).arg(&lib);
  for f in ["elmat.cu", "kuhn.cu", "assemble.cu", "tile.cu", "blockop.cu", "matfree.cu", "quadform.cu", "spmv.cu", "blas1.cu", "krylov.cu", "capi.cu"] {
    cmd.arg(csrc.join(f));
    println!("cargo:rerun-if-changed={}", csrc.join(f).display());
  }
  assert!(cmd.status().unwrap().success(), "nvcc failed");
  println!("cargo:rustc-link-search=native={}", out.display());
  println!("cargo:rustc-link-lib=dylib=formoniq_b200");
}
