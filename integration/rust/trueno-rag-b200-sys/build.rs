// Links libtrueno_rag_b200.so.  TRUENO_RAG_B200_LIB_DIR = directory that holds the library (the repository's
// trueno_rag_b200/ after `python -c "import __graft_entry__ as g; g.build()"`).
fn main() {
    println!("cargo:rerun-if-env-changed=TRUENO_RAG_B200_LIB_DIR");
    let dir = std::env::var("TRUENO_RAG_B200_LIB_DIR")
        .expect("set TRUENO_RAG_B200_LIB_DIR to the directory holding libtrueno_rag_b200.so");
    println!("cargo:rustc-link-search=native={dir}");
    println!("cargo:rustc-link-lib=dylib=trueno_rag_b200");
    // the library resolves libcudart / libcuda itself (they are DT_NEEDED entries of the .so)
    println!("cargo:rustc-link-arg=-Wl,-rpath,{dir}");
}
