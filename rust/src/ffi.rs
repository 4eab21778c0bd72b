//! FFI binding of `libsla_b200.so` (C ABI: `include/sla.h`) for the `sparse_linear_assignment` crate: `src/ffi.rs`.
//!
//! Every function `include/sla.h` declares is declared here (tests/test_host_api.py checks the two lists against each
//! other).  Not compiled in this repository -- the build image has no Rust toolchain; INTEGRATION.md lists the files a
//! maintainer adds and the patches to apply.
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct sla_ctx {
    _private: [u8; 0],
}

/// `sla_stats` of include/sla.h, field for field.
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
    pub cluster_rounds: u64,
    pub restarts: u32,
    pub reserved_: u32,
}

/// `sla_round_profile` of include/sla.h.
#[repr(C)]
#[derive(Debug, Default, Clone, Copy)]
pub struct sla_round_profile {
    pub round: u32,
    pub engine: u32,
    pub bidders: u32,
    pub rounds_covered: u32,
    pub arcs: u64,
    pub bid_ms: f32,
    pub assign_ms: f32,
}

pub const SLA_OK: c_int = 0;
pub const SLA_ERR_INVALID: c_int = 1;
pub const SLA_ERR_CUDA: c_int = 2;
pub const SLA_ERR_NO_DEVICE: c_int = 3;
pub const SLA_ERR_STATE: c_int = 4;
pub const SLA_ERR_ALLOC: c_int = 5;
pub const SLA_NONE: u32 = 0xFFFF_FFFF;
pub const SLA_ALGO_KHOSLA: c_int = 0;
pub const SLA_ALGO_FORWARD: c_int = 1;

extern "C" {
    pub fn sla_ctx_create(device: c_int, row_capacity: usize, col_capacity: usize, arc_capacity: usize,
                          out: *mut *mut sla_ctx) -> c_int;
    pub fn sla_ctx_destroy(ctx: *mut sla_ctx);
    pub fn sla_last_error(ctx: *const sla_ctx) -> *const c_char;
    pub fn sla_ctx_stream(ctx: *mut sla_ctx) -> *mut c_void;
    pub fn sla_ctx_device(ctx: *const sla_ctx) -> c_int;
    pub fn sla_version() -> *const c_char;
    pub fn sla_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn sla_host_free(p: *mut c_void);
    pub fn sla_host_negate_f64(values: *mut f64, n: usize, threads: c_int);
    pub fn sla_set_option(ctx: *mut sla_ctx, key: *const c_char, value: i64) -> c_int;
    pub fn sla_upload_csr(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, row_ptr: *const u32,
                          column_indices: *const u32, values: *const f64, nnz: u64) -> c_int;
    pub fn sla_upload_csr_negating(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, row_ptr: *const u32,
                                   column_indices: *const u32, values: *mut f64, nnz: u64, threads: c_int) -> c_int;
    pub fn sla_last_upload(ctx: *const sla_ctx, bytes: *mut u64, value_bytes: *mut u32) -> c_int;
    pub fn sla_scan_value_bytes(ctx: *const sla_ctx, value_bytes: *mut u32) -> c_int;
    pub fn sla_host_narrow(values: *mut f64, n: usize, tier: c_int, out: *mut c_void, negate: c_int) -> c_int;
    pub fn sla_upload_csr_device(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, d_row_ptr: *const u32,
                                 d_column_indices: *const u32, d_values: *const f64, nnz: u64) -> c_int;
    pub fn sla_generate_device(ctx: *mut sla_ctx, num_rows: u32, num_cols: u32, k: u32, seed: u64, value_lo: u32,
                               value_hi: u32, planted: c_int) -> c_int;
    pub fn sla_generate_device_shard(ctx: *mut sla_ctx, global_rows: u32, num_cols: u32, k: u32, seed: u64,
                                     value_lo: u32, value_hi: u32, planted: c_int, row_begin: u32, row_count: u32) -> c_int;
    pub fn sla_generate_host(num_rows: u32, num_cols: u32, k: u32, seed: u64, value_lo: u32, value_hi: u32,
                             planted: c_int, row_ptr: *mut u32, column_indices: *mut u32, values: *mut f64) -> c_int;
    pub fn sla_generate_host_ex(global_rows: u32, num_cols: u32, k: u32, seed: u64, value_lo: u32, value_hi: u32,
                                planted: c_int, value_dist: c_int, row_begin: u32, row_count: u32, threads: c_int,
                                row_ptr: *mut u32, column_indices: *mut u32, values: *mut f64) -> c_int;
    pub fn sla_khosla_solve(ctx: *mut sla_ctx, maximize: c_int, eps: f64, person_to_object: *mut u32,
                            object_to_person: *mut u32, prices: *mut f64, stats: *mut sla_stats) -> c_int;
    pub fn sla_forward_solve(ctx: *mut sla_ctx, maximize: c_int, eps: f64, start_eps: f64, max_iterations: u32,
                             person_to_object: *mut u32, object_to_person: *mut u32, prices: *mut f64,
                             stats: *mut sla_stats) -> c_int;
    pub fn sla_download_solution(ctx: *mut sla_ctx, person_to_object: *mut u32, object_to_person: *mut u32,
                                 prices: *mut f64) -> c_int;
    pub fn sla_get_objective(ctx: *mut sla_ctx, objective: *mut f64) -> c_int;
    pub fn sla_ecs_satisfied(ctx: *mut sla_ctx, eps: f64, toleration: f64, satisfied: *mut c_int) -> c_int;
    pub fn sla_validate_matching(ctx: *mut sla_ctx, num_unassigned: *mut u32, consistent: *mut c_int) -> c_int;
    pub fn sla_get_round_profile(ctx: *mut sla_ctx, out: *mut sla_round_profile, capacity: usize, count: *mut usize) -> c_int;
    pub fn sla_batch_upload(ctx: *mut sla_ctx, num_instances: u32, row_off: *const u32, col_off: *const u32,
                            row_ptr: *const u32, column_indices: *const u32, values: *const f64) -> c_int;
    pub fn sla_batch_generate_device(ctx: *mut sla_ctx, num_instances: u32, first_instance_id: u32, num_rows: u32,
                                     num_cols: u32, k: u32, seed: u64, value_lo: u32, value_hi: u32, planted: c_int) -> c_int;
    pub fn sla_batch_solve(ctx: *mut sla_ctx, algo: c_int, maximize: c_int, eps: f64, start_eps: f64,
                           max_iterations: u32, person_to_object: *mut u32, object_to_person: *mut u32,
                           prices: *mut f64, per_instance_stats: *mut sla_stats, total: *mut sla_stats) -> c_int;
    pub fn sla_part_local_value_range(ctx: *mut sla_ctx, w_min: *mut f64, w_max: *mut f64, first_value: *mut f64) -> c_int;
    pub fn sla_part_begin(ctx: *mut sla_ctx, algo: c_int, maximize: c_int, row_begin: u32, global_rows: u32,
                          eps: f64, global_w_min: f64, global_w_max: f64, global_first_value: f64) -> c_int;
    pub fn sla_part_buffers(ctx: *mut sla_ctx, d_best_words: *mut *mut c_void,
                            d_price_candidates: *mut *mut c_void, num_words: *mut u64) -> c_int;
    pub fn sla_part_bid(ctx: *mut sla_ctx) -> c_int;
    pub fn sla_part_claim(ctx: *mut sla_ctx) -> c_int;
    pub fn sla_part_assign(ctx: *mut sla_ctx, local_queue_len: *mut u32, local_dropped: *mut u32) -> c_int;
    pub fn sla_part_finish(ctx: *mut sla_ctx, person_to_object: *mut u32, object_to_person: *mut u32,
                           prices: *mut f64, stats: *mut sla_stats) -> c_int;
    pub fn sla_part_sparse_buffers(ctx: *mut sla_ctx, world: c_int, d_send: *mut *mut c_void,
                                   d_recv: *mut *mut c_void, d_counts: *mut *mut c_void, send_capacity: *mut u64) -> c_int;
    pub fn sla_part_collect(ctx: *mut sla_ctx, local_winners: *mut u32) -> c_int;
    pub fn sla_part_apply_sparse(ctx: *mut sla_ctx, world: c_int, max_count: u32, local_queue_len: *mut u32,
                                 local_dropped: *mut u32) -> c_int;
    pub fn sla_mesh_create(ctx: *mut sla_ctx, rank: c_int, world: c_int, row_begins: *const u32, global_cols: u32,
                           block: *mut *mut c_void, block_bytes: *mut usize) -> c_int;
    pub fn sla_ipc_export(dev_ptr: *mut c_void, handle64: *mut u8) -> c_int;
    pub fn sla_ipc_import(device: c_int, handle64: *const u8, dev_ptr: *mut *mut c_void) -> c_int;
    pub fn sla_ipc_release(device: c_int, dev_ptr: *mut c_void) -> c_int;
    pub fn sla_mesh_connect(ctx: *mut sla_ctx, peer_blocks: *const *mut c_void, peer_devices: *const c_int) -> c_int;
    pub fn sla_mesh_begin(ctx: *mut sla_ctx, maximize: c_int, eps: f64, global_w_min: f64, global_w_max: f64,
                          global_first_value: f64) -> c_int;
    pub fn sla_mesh_solve(ctx: *mut sla_ctx) -> c_int;
    pub fn sla_mesh_phase(ctx: *mut sla_ctx, which: c_int) -> c_int;
    pub fn sla_mesh_poll(ctx: *mut sla_ctx, done: *mut c_int, round: *mut u32, local_queue_len: *mut u32) -> c_int;
    pub fn sla_mesh_finish(ctx: *mut sla_ctx, person_to_object: *mut u32, object_to_person: *mut u32,
                           prices: *mut f64, stats: *mut sla_stats) -> c_int;
    pub fn sla_mesh_owned(ctx: *mut sla_ctx, shard_objects: *mut u32, num_owned: *mut u32, first_object: *mut u32,
                          first_row: *mut u32) -> c_int;
    pub fn sla_mesh_round1_ms(ctx: *mut sla_ctx, bid_ms: *mut f32) -> c_int;
    pub fn sla_mesh_objective(ctx: *mut sla_ctx, objective: *mut f64) -> c_int;
    pub fn sla_mesh_timeline(ctx: *mut sla_ctx, out: *mut u64, capacity: usize) -> c_int;
}
