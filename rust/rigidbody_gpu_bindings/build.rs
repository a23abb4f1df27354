// Links the prebuilt CUDA library (make -C rigidbody_rs_b200/csrc EXTRA_NVFLAGS=-DRB_NO_REFERENCE_SYMBOLS).
// RIGIDBODY_B200_LIB_DIR must point at the directory holding librigidbody_b200.so.
fn main() {
    let dir = std::env::var("RIGIDBODY_B200_LIB_DIR").expect("set RIGIDBODY_B200_LIB_DIR");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=rigidbody_b200");
    println!("cargo:rerun-if-env-changed=RIGIDBODY_B200_LIB_DIR");
}
