// UNVERIFIED SOURCE (no rustc in the build image).
// Replaces hpt-cudakernels/build.rs: nothing is compiled here; libhpt_b200.so is produced by `python build.py`
// (nvcc -gencode arch=compute_100a,code=sm_100a).  HPT_B200_LIB_DIR points at hpt_b200/lib.
fn main() {
    let dir = std::env::var("HPT_B200_LIB_DIR").unwrap_or_else(|_| "../../hpt_b200/lib".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=hpt_b200");
    println!("cargo:rerun-if-env-changed=HPT_B200_LIB_DIR");
}
