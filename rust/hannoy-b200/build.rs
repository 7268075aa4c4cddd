// Links libhannoy_b200.so (built by `python -m hannoy_b200.build`, nvcc sm_100a).
// HANNOY_B200_LIB_DIR = directory that holds the shared library.
fn main() {
    if let Ok(dir) = std::env::var("HANNOY_B200_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=hannoy_b200");
    println!("cargo:rerun-if-env-changed=HANNOY_B200_LIB_DIR");
}
