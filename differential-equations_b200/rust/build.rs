// Link against the in-tree shared library built by `python __graft_entry__.py` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("DEB200_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/..", std::env::var("CARGO_MANIFEST_DIR").unwrap())
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=deb200");
    println!("cargo:rerun-if-env-changed=DEB200_LIB_DIR");
}
