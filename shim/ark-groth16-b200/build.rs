// Links against groth16_b200/libg16b200.so (built by `make -C groth16_b200/csrc`).  G16B200_LIB_DIR overrides the path.
fn main() {
    let dir = std::env::var("G16B200_LIB_DIR").unwrap_or_else(|_| {
        let here = std::path::PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap());
        here.join("../../groth16_b200").to_string_lossy().into_owned()
    });
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=g16b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    println!("cargo:rerun-if-env-changed=G16B200_LIB_DIR");
}
