// Links librtgpu.so (the hand-written sm_100a kernels + the C ABI of include/rtgpu.h).
//
//   default feature set   : RTGPU_LIB_DIR=/path/to/ray_tracer_challenge_rs_b200 (the directory that holds librtgpu.so)
//   build-from-source     : RTGPU_SRC_DIR=/path/to/the/rtgpu/repository; needs nvcc (CUDA >= 12.8 for sm_100a)
//
// -fmad=false is part of the numerics contract: the reference never contracts a*b+c, and the kernels call fma()
// exactly where the reference calls mul_add.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=RTGPU_LIB_DIR");
    println!("cargo:rerun-if-env-changed=RTGPU_SRC_DIR");

    #[cfg(feature = "build-from-source")]
    {
        let src = env::var("RTGPU_SRC_DIR").expect("build-from-source needs RTGPU_SRC_DIR (root of the rtgpu repository)");
        cc::Build::new()
            .cuda(true)
            .cudart("shared")
            .flag("-gencode")
            .flag("arch=compute_100a,code=sm_100a")
            .flag("-O3")
            .flag("-std=c++17")
            .flag("-lineinfo")
            .flag("-fmad=false")
            .flag("-prec-div=true")
            .flag("-prec-sqrt=true")
            .flag("-ftz=false")
            .include(format!("{src}/include"))
            .file(format!("{src}/ray_tracer_challenge_rs_b200/csrc/rtgpu.cu"))
            .compile("rtgpu");
        println!("cargo:rustc-link-lib=dylib=cudart");
        println!("cargo:rustc-link-lib=dylib=stdc++");
        return;
    }

    #[cfg(not(feature = "build-from-source"))]
    {
        let dir = env::var("RTGPU_LIB_DIR")
            .expect("set RTGPU_LIB_DIR to the directory that holds librtgpu.so (or enable the build-from-source feature)");
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-lib=dylib=rtgpu");
        // so that `cargo run` finds the library without LD_LIBRARY_PATH
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
}
