// Links libbsx.so (built by `make -C blobstreamx_b200/csrc`).  BSX_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("BSX_LIB_DIR").unwrap_or_else(|_| "../../../blobstreamx_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bsx");
    println!("cargo:rerun-if-env-changed=BSX_LIB_DIR");
}
