// Link against libp2gpu.so built by `python -c "import __graft_entry__ as g; g.build()"`.
fn main() {
    let dir = std::env::var("P2GPU_LIB_DIR").expect("set P2GPU_LIB_DIR to <repo>/plonky2_aes_b200");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=p2gpu");
    println!("cargo:rerun-if-env-changed=P2GPU_LIB_DIR");
}
