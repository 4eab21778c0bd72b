//! FFI binding of `libsla_b200.so` (C ABI: `include/sla.h`) for the `sparse_linear_assignment` crate.
//!
//! Not compiled in this repository (no Rust toolchain in the build image); it is the binding a maintainer adds as
//! `src/ffi.rs`, together with `build.rs` (`println!("cargo:rustc-link-lib=dylib=sla_b200")`) and a `links = "sla_b200"`
//! line in `Cargo.toml`.  See INTEGRATION.md for the two `solve` bodies that call it.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct sla_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct sla_stats {
    pub num_unassigned: u32,
    pub nits: u32,
    pub nreductions: u32,
    pub optimal_soln_found: u32,
    pub eps: f64,
    pub rounds: u64,
    pub bids: u64,
    pub bid_arcs: u64,
    pub dropped: u32,
    pub values_negated: u32,
    pub wide_rounds: u64,
    pub tail_rounds: u64,
    pub kernel_launches: u32,
    pub graph_launches: u32,
    pub ms_solve: f32,
    pub ms_total: f32,
}

pub const SLA_OK: c_int = 0;

extern "C" {
    pub fn sla_ctx_create(device: c_int, row_capacity: usize, col_capacity: usize, arc_capacity: usize,
                          out: *mut *mut sla_ctx) -> c_int;
    pub fn sla_ctx_destroy(ctx: *mut sla_ctx);
    pub fn sla_last_error(ctx: *const sla_ctx) -> *const c_char;
    pub fn sla_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn sla_host_free(p: *mut c_void);
    pub fn sla_host_negate_f64(values: *mut f64, n: usize, threads: c_int);
    pub fn sla_upload_csr(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, row_ptr: *const u32,
                          column_indices: *const u32, values: *const f64, nnz: u64) -> c_int;
    pub fn sla_upload_csr_negating(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, row_ptr: *const u32,
                                   column_indices: *const u32, values: *mut f64, nnz: u64, threads: c_int) -> c_int;
    pub fn sla_host_narrow(values: *mut f64, n: usize, tier: c_int, out: *mut c_void, negate: c_int) -> c_int;
    pub fn sla_last_upload(ctx: *const sla_ctx, bytes: *mut u64, value_bytes: *mut u32) -> c_int;
    pub fn sla_scan_value_bytes(ctx: *const sla_ctx, value_bytes: *mut u32) -> c_int;
    pub fn sla_khosla_solve(ctx: *mut sla_ctx, maximize: c_int, eps: f64, person_to_object: *mut u32,
                            object_to_person: *mut u32, prices: *mut f64, stats: *mut sla_stats) -> c_int;
    pub fn sla_forward_solve(ctx: *mut sla_ctx, maximize: c_int, eps: f64, start_eps: f64, max_iterations: u32,
                             person_to_object: *mut u32, object_to_person: *mut u32, prices: *mut f64,
                             stats: *mut sla_stats) -> c_int;
    pub fn sla_download_solution(ctx: *mut sla_ctx, person_to_object: *mut u32, object_to_person: *mut u32,
                                 prices: *mut f64) -> c_int;
    pub fn sla_get_objective(ctx: *mut sla_ctx, objective: *mut f64) -> c_int;
    pub fn sla_ecs_satisfied(ctx: *mut sla_ctx, eps: f64, toleration: f64, satisfied: *mut c_int) -> c_int;
}

/// Owns one device context; stored as a field of `KhoslaSolver` / `ForwardAuctionSolver`.
/// `Clone` creates a fresh context lazily (the CSR is re-uploaded on the clone's first solve).
pub struct DeviceMirror {
    pub ctx: *mut sla_ctx,
    pub dirty: bool,
}

impl DeviceMirror {
    pub fn new(rows: usize, cols: usize, arcs: usize) -> anyhow::Result<Self> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { sla_ctx_create(0, rows, cols, arcs, &mut ctx) };
        if rc != SLA_OK {
            let msg = unsafe { std::ffi::CStr::from_ptr(sla_last_error(std::ptr::null())) };
            anyhow::bail!("sla_ctx_create failed ({}): {}", rc, msg.to_string_lossy());
        }
        Ok(Self { ctx, dirty: true })
    }

    pub fn check(&self, rc: c_int) -> anyhow::Result<()> {
        if rc == SLA_OK {
            return Ok(());
        }
        let msg = unsafe { std::ffi::CStr::from_ptr(sla_last_error(self.ctx)) };
        Err(anyhow::anyhow!("libsla_b200 error {}: {}", rc, msg.to_string_lossy()))
    }
}

impl Drop for DeviceMirror {
    fn drop(&mut self) {
        unsafe { sla_ctx_destroy(self.ctx) }
    }
}

unsafe impl Send for DeviceMirror {}
