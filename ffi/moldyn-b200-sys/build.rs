// Links libmoldyn_b200.so (built in-tree by `python -m moldyn_b200.build`, nvcc for sm_100a).
fn main() {
    let dir = std::env::var("MOLDYN_B200_LIB_DIR")
        .expect("set MOLDYN_B200_LIB_DIR to <moldyn_b200 checkout>/moldyn_b200/lib");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=moldyn_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=MOLDYN_B200_LIB_DIR");
}
