// build.rs of the patched crate: link against libsla_b200.so (built by `python -c "import __graft_entry__ as g; g.build()"`
// in the sparse_linear_assignment_b200 repository).  SLA_B200_LIB_DIR = the directory that holds the shared object.
fn main() {
    let dir = std::env::var("SLA_B200_LIB_DIR").expect("set SLA_B200_LIB_DIR to the directory that contains libsla_b200.so");
    println!("cargo:rustc-link-search=native={}", dir);
    println!("cargo:rustc-link-lib=dylib=sla_b200");
    println!("cargo:rustc-link-arg=-Wl,-rpath,{}", dir);
    println!("cargo:rerun-if-env-changed=SLA_B200_LIB_DIR");
}
