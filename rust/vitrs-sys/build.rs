// Links the prebuilt C-ABI library.  Build it first: `make -C vit.rs_b200/csrc` (nvcc, sm_100a).
fn main() {
    let dir = std::env::var("VITRS_LIB_DIR").unwrap_or_else(|_| "../../vit.rs_b200".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=vitrs");
    println!("cargo:rerun-if-env-changed=VITRS_LIB_DIR");
}
