// Link against libbsr_b200.so built by `make -C better_search_rag_rust_b200/csrc`.
fn main() {
    let dir = std::env::var("BSR_B200_LIB_DIR")
        .unwrap_or_else(|_| "../better_search_rag_rust_b200".to_string());
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=bsr_b200");
}
