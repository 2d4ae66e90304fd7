// Links the in-tree shared library.  RT_B200_LIB_DIR points at cpu-raytracing-rt_b200/ (where
// `make -C cpu-raytracing-rt_b200/csrc` leaves librt_b200.so).
fn main() {
    let dir = std::env::var("RT_B200_LIB_DIR").unwrap_or_else(|_| "../../cpu-raytracing-rt_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=rt_b200");
    println!("cargo:rerun-if-env-changed=RT_B200_LIB_DIR");
}
