//! Addition to the crate's `src/build.rs` (which today only links the Windows resource file,
//! build.rs:1-39): find libaptb200.so and, optionally, regenerate the bindings.  SOURCE ONLY.
fn main() {
    // directory holding libaptb200.so (built by `python -c "import __graft_entry__ as g; g.build()"`)
    let dir = std::env::var("APTB200_LIB_DIR").unwrap_or_else(|_| "../noaa-apt_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=aptb200");
    println!("cargo:rerun-if-env-changed=APTB200_LIB_DIR");
    // With the `bindgen` build-dependency the hand-written rust/aptb200_sys.rs can be replaced by:
    //   bindgen::Builder::default().header("../include/aptb200.h").generate().unwrap()
    //       .write_to_file(std::path::Path::new(&std::env::var("OUT_DIR").unwrap()).join("aptb200_sys.rs")).unwrap();
}
