// Links the prebuilt CUDA library.  SIGOPS_LIB_DIR defaults to ../wgpu-sigops_b200 (where build.py puts libsigops.so).
fn main() {
    let dir = std::env::var("SIGOPS_LIB_DIR").unwrap_or_else(|_| {
        format!("{}/../wgpu-sigops_b200", std::env::var("CARGO_MANIFEST_DIR").unwrap())
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=sigops");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=SIGOPS_LIB_DIR");
}
