// Links libbpp_b200.so (built by `make -C bulletproofs-plus_b200/csrc`, or `python -c "import __graft_entry__ as g; g.build()"`).
use std::env;
use std::path::PathBuf;

fn main() {
    let dir = env::var("BPP_B200_LIB_DIR").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../bulletproofs-plus_b200")
    });
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=bpp_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
    println!("cargo:rerun-if-env-changed=BPP_B200_LIB_DIR");
}
