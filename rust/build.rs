// Links libssw.so (built by `make -C spread_spectrum_watermarking_b200/csrc`).
// SSW_LIB_DIR overrides the default in-tree location.
fn main() {
    let dir = std::env::var("SSW_LIB_DIR")
        .unwrap_or_else(|_| format!("{}/../spread_spectrum_watermarking_b200/csrc", env!("CARGO_MANIFEST_DIR")));
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=ssw");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=SSW_LIB_DIR");
}
