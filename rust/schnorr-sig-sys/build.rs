// Links against the in-tree build of the CUDA library (python schnorr-sig_b200/build.py writes
// schnorr-sig_b200/csrc/libschnorr_b200.so).  SCHNORR_B200_LIB_DIR overrides the search directory at BUILD time.
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("SCHNORR_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../schnorr-sig_b200/csrc")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=schnorr_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=SCHNORR_B200_LIB_DIR");
    println!("cargo:rerun-if-changed=../../include/schnorr_b200.h");
}
