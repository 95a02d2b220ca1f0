// Links the prebuilt CUDA library; EAGEN_MSM_LIB_DIR points at halo2-liam-eagen-msm_b200/ (where build.py leaves libeagen_msm.so).
fn main() {
    let dir = std::env::var("EAGEN_MSM_LIB_DIR").unwrap_or_else(|_| "..".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=eagen_msm");
    println!("cargo:rerun-if-env-changed=EAGEN_MSM_LIB_DIR");
}
