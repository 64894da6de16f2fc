// Links libgvdb.so; GVDB_LIB_DIR points at grape-vector-db_b200/lib.
fn main() {
    let dir = std::env::var("GVDB_LIB_DIR").unwrap_or_else(|_| "../lib".to_string());
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=gvdb");
    println!("cargo:rerun-if-env-changed=GVDB_LIB_DIR");
}
