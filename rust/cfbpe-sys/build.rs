// Links libcfbpe.so (built by `python -c 'import __graft_entry__ as g; g.build()'` in the tokenizer repository, or shipped in
// the plugin's container image).  CFBPE_LIB_DIR names the directory that holds it.
fn main() {
    if let Ok(dir) = std::env::var("CFBPE_LIB_DIR") {
        println!("cargo:rustc-link-search=native={dir}");
        println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
    }
    println!("cargo:rustc-link-lib=dylib=cfbpe");
    println!("cargo:rerun-if-env-changed=CFBPE_LIB_DIR");
}
