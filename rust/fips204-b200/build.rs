// Links the C-ABI library of the engine.  MLDSA_B200_LIB_DIR = directory that holds libmldsa_b200.so (default: the
// in-tree build, ../../fips204_b200, produced by `python -m fips204_b200.build`).
use std::{env, path::PathBuf};

fn main() {
    let dir = env::var("MLDSA_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../fips204_b200")
    });
    println!("cargo:rerun-if-env-changed=MLDSA_B200_LIB_DIR");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=mldsa_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
