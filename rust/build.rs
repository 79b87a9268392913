// build.rs of the decaf377 crate when it is built with `--features b200`: tells cargo where
// libdecaf377_b200.so lives (the in-tree build product of `python -m decaf377_b200.build`).
//
//   D377_B200_LIB_DIR=/path/to/decaf377_b200 cargo build --features b200
//
// The `#[link(name = "decaf377_b200")]` attribute in src/gpu.rs names the library; this
// script only adds the search path and an rpath so that tests and binaries find it at run
// time without LD_LIBRARY_PATH.
use std::env;

fn main() {
    println!("cargo:rerun-if-env-changed=D377_B200_LIB_DIR");
    if env::var_os("CARGO_FEATURE_B200").is_none() {
        return;
    }
    let dir = env::var("D377_B200_LIB_DIR")
        .expect("--features b200 needs D377_B200_LIB_DIR: the directory that holds libdecaf377_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=decaf377_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
}
