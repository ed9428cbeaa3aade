// Links libllkv_gpu.so (built by rust-llkv_b200/csrc/build.sh).  LLKV_GPU_LIB_DIR points at the directory holding it.
fn main() {
    let dir = std::env::var("LLKV_GPU_LIB_DIR").unwrap_or_else(|_| "../../rust-llkv_b200/csrc".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=llkv_gpu");
    println!("cargo:rerun-if-env-changed=LLKV_GPU_LIB_DIR");
}
