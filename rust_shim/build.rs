// Links the in-tree CUDA library (python -m fhe_regex_b200.build produces fhe_regex_b200/libfhe_b200.so).
fn main() {
    let dir = std::env::var("FHE_B200_LIB_DIR").unwrap_or_else(|_| "../fhe_regex_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=fhe_b200");
    println!("cargo:rerun-if-env-changed=FHE_B200_LIB_DIR");
}
