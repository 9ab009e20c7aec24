// Links libpfhe_cuda.so (built by `python -m primus_fhe_b200.build`, nvcc -gencode arch=compute_100a,code=sm_100a).
// PFHE_LIB_DIR points at <repo>/primus_fhe_b200/lib; the library links the CUDA runtime statically and needs only libcuda.so.1.
fn main() {
    let dir = std::env::var("PFHE_LIB_DIR").expect("set PFHE_LIB_DIR to the directory that holds libpfhe_cuda.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=pfhe_cuda");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=PFHE_LIB_DIR");
}
