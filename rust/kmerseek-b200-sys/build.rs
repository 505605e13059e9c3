// NOT COMPILED HERE (no Rust toolchain in the build image).
// Links the prebuilt shared library; set KMERSEEK_B200_LIB_DIR to the directory that holds libkmerseek_b200.so
// (kmerseek_b200/ in this repository after `python -m kmerseek_b200.build`).
fn main() {
    let dir = std::env::var("KMERSEEK_B200_LIB_DIR").expect("set KMERSEEK_B200_LIB_DIR");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=kmerseek_b200");
    println!("cargo:rerun-if-changed=../../include/kmerseek_b200.h");
    // With bindgen available the block in src/lib.rs can be generated instead:
    //   bindgen::Builder::default().header("../../include/kmerseek_b200.h").allowlist_function("ks_.*") ...
}
