// Links liblight_garden_b200.so (built in-tree by `python -c 'import __graft_entry__ as g; g.build()'` into
// light_garden_b200/_lib/).  LIGHT_GARDEN_B200_LIB overrides the directory; the default is the in-tree location
// relative to this crate (rust/lightgarden-cuda-sys -> ../../light_garden_b200/_lib).
use std::path::PathBuf;

fn main() {
    let dir = std::env::var("LIGHT_GARDEN_B200_LIB").map(PathBuf::from).unwrap_or_else(|_| {
        PathBuf::from(std::env::var("CARGO_MANIFEST_DIR").unwrap()).join("../../light_garden_b200/_lib")
    });
    println!("cargo:rerun-if-env-changed=LIGHT_GARDEN_B200_LIB");
    println!("cargo:rustc-link-search=native={}", dir.display());
    println!("cargo:rustc-link-lib=dylib=light_garden_b200");
    // so that `cargo run` finds the shared object without LD_LIBRARY_PATH
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir.display());
}
