//! Links libdllm_b200.so.  Default: the library built by `make -C diffusion-llm-rs_b200` (found through DLLM_B200_ROOT, or
//! two directories above this crate).  With `--features build-cuda`, `cc` drives nvcc over csrc/*.cu for sm_100a itself.
//! There is no CPU fallback: without the library (and, at run time, an sm_100 GPU) nothing links / dllm_ctx_create fails.
use std::{env, path::PathBuf};

fn main() {
    let root = env::var("DLLM_B200_ROOT")
        .map(PathBuf::from)
        .unwrap_or_else(|_| PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../.."));
    let pkg = root.join("diffusion-llm-rs_b200");
    println!("cargo:rerun-if-env-changed=DLLM_B200_ROOT");
    println!("cargo:rerun-if-changed={}", root.join("include/dllm_b200.h").display());

    #[cfg(feature = "build-cuda")]
    {
        let mut b = cc::Build::new();
        b.cuda(true)
            .flag("-gencode")
            .flag("arch=compute_100a,code=sm_100a")
            .flag("-std=c++17")
            .flag("-O3")
            .flag("-lineinfo")
            .flag("-fmad=false") // the quantizer arithmetic must not be contracted into FMAs (Rust never contracts)
            .flag("--expt-relaxed-constexpr");
        for f in ["api", "quant_kernels", "weight_kernels", "gemv_simt", "gemv_mma", "umma_gemm", "sample_kernels", "tp"] {
            let p = pkg.join("csrc").join(format!("{f}.cu"));
            println!("cargo:rerun-if-changed={}", p.display());
            b.file(p);
        }
        b.compile("dllm_b200");
        println!("cargo:rustc-link-lib=dylib=cudart");
        println!("cargo:rustc-link-lib=dylib=nccl");
        return;
    }

    #[cfg(not(feature = "build-cuda"))]
    {
        println!("cargo:rustc-link-search=native={}", pkg.join("lib").display());
        println!("cargo:rustc-link-lib=dylib=dllm_b200");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{}", pkg.join("lib").display());
    }
}
