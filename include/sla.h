/*
 * sla.h -- C ABI of the B200-native auction hot path (libsla_b200.so).
 *
 * This is the drop-in boundary for DXist/sparse_linear_assignment v0.1.5: the Rust crate keeps its public
 * API and its host-side CSR storage, and the bodies of
 *     KhoslaSolver::solve                       (reference src/ksparse.rs:153-251)
 *     ForwardAuctionSolver::solve_with_params   (reference src/symmetric.rs:217-332, bid_and_assign 334-468,
 *                                                push_all_left 471-508)
 * call the entry points below over FFI (binding shown in INTEGRATION.md, rust/ffi.rs).  The reference has no
 * FFI of its own, so every entry point cites the reference interface it stands in for.
 *
 * Conventions: plain pointers and sizes only; every function returns an int status (SLA_OK == 0); no
 * exceptions cross the boundary; `sla_last_error` returns a message for the last failure.  Index type on the
 * boundary is uint32_t (the reference's u16 instantiation is widened by the host wrapper; the unassigned
 * sentinel SLA_NONE == u32::MAX truncates to u16::MAX).  A context is not thread-safe (one CUDA stream per
 * context, like `&mut self` in the reference); distinct contexts may be used from distinct threads.
 * There is no CPU fallback: every call fails with SLA_ERR_NO_DEVICE when no sm_100 GPU is usable.
 */
#ifndef SLA_B200_H
#define SLA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLA_OK 0
#define SLA_ERR_INVALID 1   /* a reference `ensure!` would fail: bad sizes, empty input, index out of range */
#define SLA_ERR_CUDA 2      /* CUDA runtime / driver error (message has the CUDA error string) */
#define SLA_ERR_NO_DEVICE 3 /* no usable GPU */
#define SLA_ERR_STATE 4     /* call order (solve before upload, ...) or safety round limit hit */
#define SLA_ERR_ALLOC 5     /* device or pinned-host allocation failed */

#define SLA_NONE 0xFFFFFFFFu /* I::MAX "unassigned" sentinel, reference src/solution.rs:27-34 */

#define SLA_ALGO_KHOSLA 0
#define SLA_ALGO_FORWARD 1

typedef struct sla_ctx sla_ctx;

/* Result scalars of one solve.  Mirrors the reference's observable post-conditions:
 *   num_unassigned, eps            -> AuctionSolution fields (src/solution.rs:35-39)
 *   nits                           -> KhoslaSolver::nits (ksparse.rs:84; bids made, dropped ones included) or
 *                                     ForwardAuctionSolver::nits (symmetric.rs:88; Jacobi rounds)
 *   nreductions, optimal_soln_found-> symmetric.rs:89-90
 *   values_negated                 -> 1 when AuctionSolver::init_solve would have negated `values` in place
 *                                     (solver.rs:209-216); the host wrapper must then negate its own copy.
 * The remaining fields are instrumentation the reference does not have. */
typedef struct sla_stats {
    uint32_t num_unassigned;
    uint32_t nits;
    uint32_t nreductions;
    uint32_t optimal_soln_found;
    double eps;
    uint64_t rounds;      /* synchronous Jacobi rounds, all phases */
    uint64_t bids;        /* sum over rounds of bidders */
    uint64_t bid_arcs;    /* sum over rounds of (column, value) pairs examined: the unit of BASELINE.json's metric */
    uint32_t dropped;     /* Khosla: persons dropped by the price threshold (ksparse.rs:218-220) */
    uint32_t values_negated;
    uint64_t wide_rounds; /* rounds run by the grid-wide kernels */
    uint64_t tail_rounds; /* rounds run by the persistent engines: the single-CTA tail engine and the cluster engine */
    uint32_t kernel_launches; /* kernels launched for this solve (graph nodes included) */
    uint32_t graph_launches;
    float ms_solve;       /* device time of the solve, CUDA events on the context's stream, D2H excluded */
    float ms_total;       /* ms_solve + result copies */
    uint64_t cluster_rounds; /* of tail_rounds: rounds run by the thread-block-cluster engine (queues of a few thousand bidders
                                on instances whose prices do not fit one CTA's shared memory) */
    uint32_t restarts;    /* Khosla on a square instance (option "khosla_scaling"): 1 when a phase of the eps-schedule dropped
                             somebody and the solve started over with the plain fixed-eps rounds; batch
                             totals: number of instances that did.  nreductions counts the phases of the schedule */
    uint32_t reserved_;   /* keeps the size a multiple of 8 on every ABI; always 0 */
} sla_stats;

/* One record per bid-scan launch when option "profile" is 1 (host-driven loop, CUDA events around each kernel). */
typedef struct sla_round_profile {
    uint32_t round;
    uint32_t engine;   /* 0 = wide kernels, 1 = single-CTA tail engine (covers `rounds_covered` rounds) */
    uint32_t bidders;
    uint32_t rounds_covered;
    uint64_t arcs;
    float bid_ms;
    float assign_ms;
} sla_round_profile;

/* ---- context: replaces the state owned by the solver structs (ksparse.rs:73-85, symmetric.rs:75-98);
 *      created from `new(row_capacity, column_capacity, arcs_capacity)` (solver.rs:9-13), freed in Drop ---- */
int sla_ctx_create(int device, size_t row_capacity, size_t col_capacity, size_t arc_capacity, sla_ctx **out);
void sla_ctx_destroy(sla_ctx *ctx);
const char *sla_last_error(const sla_ctx *ctx); /* ctx may be NULL: error of the last failed sla_ctx_create */
void *sla_ctx_stream(sla_ctx *ctx);             /* the context's cudaStream_t, for event timing by the caller */
int sla_ctx_device(const sla_ctx *ctx);
const char *sla_version(void);

/* Host-memory helpers for the wrappers: page-locked staging buffers (a Rust shim would back its Vecs with these, or
 * cudaHostRegister them) and the in-place negation of the host copy of `values` that AuctionSolver::init_solve
 * performs (solver.rs:214-216; stats.values_negated says when), split over `threads` host threads. */
int sla_host_alloc(size_t bytes, void **out);
void sla_host_free(void *p);
void sla_host_negate_f64(double *values, size_t n, int threads);

/* Options: "tail_max" (bidders at or below which the single-CTA tail engine runs, <= 1024; the engine may lower it
 * to make room for its shared-memory mirrors), "smem_prices" / "smem_owners" (1: let the tail engine mirror object
 * prices / owners in shared memory when they fit), "graph" (1: CUDA-graph super-rounds, 0: host-driven loop),
 * "zero_price_skip" (1: skip the price gather while all prices are exactly 0, i.e. the first round after
 * init_solve), "regular" (1: use the uniform-degree bid kernel when every row has the same multiple-of-8 degree),
 * "khosla_scaling" (1: Khosla rounds on square instances run under an eps-schedule that ends at the caller's eps
 * and fall back to the plain rounds when a phase drops anybody; 0: plain fixed-eps rounds),
 * "profile" (1: record sla_round_profile entries), "super_rounds" (rounds captured per graph), "timeout_s",
 * "small_path" (1: instances of up to 1 MB take their statistics on the host while they are staged for upload and
 * download their results behind each graph launch: one device round trip per solve), "wide_first" (1: plain Khosla
 * solves with 641 .. tail_max persons run their first round on the grid-wide kernels), "narrow_upload" (1: see
 * sla_last_upload), "narrow_scan" (1: see sla_scan_value_bytes), "stream_scan" (1: first-round scan of a uniform-degree CSR through the TMA pipeline
 * bid_stream_kernel instead of the LDG.256 kernel; identical results), "l2_persist" (1: access-policy window that
 * keeps the bid words persisting in the L2), "profile_repeat" (development: launches of the scan per profile bracket),
 * "prune_gather" (1: rounds with >= 32 Ki bidders gather only the prices that can still matter -- prices never fall, so
 * profit <= value; identical choices), "learn_shape" (1: the first graph of a plain Khosla solve of a resident CSR is
 * shaped by the previous solve's number of wide rounds), "upload_ride" (1: u16 uploads run their statistics and the
 * widening behind every run of staged pieces on a second stream instead of two passes at the end), "cluster_engine"
 * (0; 1: queues of 25 .. 8,192 bidders of large-M instances run in one thread-block cluster, mid_kernel; identical
 * results, measured no faster) with "cluster_handover" (queue length at which it hands over to the single-CTA engine),
 * "mesh_tail" (1: the mesh engine's persistent one-block tail), "prezero_best" (0; development), "profile_graph"
 * (1: the profile bracket is one small graph launch).  Every option is a performance / engine-selection switch: none
 * changes a result bit, with ONE documented exception in what is compared against:
 *
 * Guarantee for KhoslaSolver on SQUARE instances (num_rows == num_cols) with "khosla_scaling" = 1 (the default): the
 * returned assignment and prices satisfy eps-complementary slackness at the caller's eps (the last phase runs at exactly
 * that eps) and num_unassigned equals the reference's -- that, not equality of prices / nits / the assignment itself
 * with the reference's single fixed-eps pass (ksparse.rs:153-251), is what is promised: optimal objective for integer
 * weights with eps < 1/n, within n * eps otherwise.  stats.nreductions then counts the phases and stats.restarts says
 * whether the schedule was abandoned for the plain rounds (a phase dropped somebody at the price threshold,
 * ksparse.rs:218-220: such instances return exactly what "khosla_scaling" = 0 returns).  Rectangular instances never
 * run a schedule. */
int sla_set_option(sla_ctx *ctx, const char *key, int64_t value);

/* ---- CSR mirror: the host keeps ownership of i_starts_stops / column_indices / values built by
 *      init / add_value / extend_from_values (solver.rs:41-101, 191-205); this copies them to HBM and
 *      computes the value statistics the solves need (min, max, max |a_ij|; ksparse.rs:171-179,
 *      symmetric.rs:246).  row_ptr has num_rows + 1 entries (row extents are (row_ptr[i], row_ptr[i+1]),
 *      equivalent to the reference's (i_starts_stops[i], j_counts[i]), solver.rs:88-97).
 *      Performs validate_input (solver.rs:232-243) plus the column bound the reference only debug_asserts. */
int sla_upload_csr(sla_ctx *ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *row_ptr,
                   const uint32_t *column_indices, const double *values, uint64_t nnz);

/* sla_upload_csr plus the in-place negation of the HOST `values` that AuctionSolver::init_solve performs
 * (solver.rs:214-216) when `maximize ^ (values[0] >= 0)`, done by up to `threads` host worker threads owned by the
 * context while the upload and the solve run: in the pass that stages a narrow (u16 / f32) copy of the values for the
 * wire when they allow it (sla_last_upload), else chunk by chunk behind the f64 copies; small instances: in the single
 * staging pass of the calling thread.  The work is complete before the next sla_*_solve / upload / destroy on this
 * context returns.  The device keeps the original values: the following solve still reports values_negated == 1,
 * and the host must not negate again. */
int sla_upload_csr_negating(sla_ctx *ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *row_ptr,
                            const uint32_t *column_indices, double *values, uint64_t nnz, int threads);

/* Bytes the last sla_upload_csr / sla_upload_csr_negating call moved from host to device, and the width (2, 4 or 8
 * bytes) the values crossed PCIe with.  Large uploads whose values all survive a round trip through u16 (non-negative
 * integers below 65,536) or f32 are staged narrow by the context's host workers and widened to f64 again in HBM, bit
 * for bit; option "narrow_upload" = 0 turns this off.  No counterpart in the reference (its Vecs never leave the
 * host, solver.rs:41-101). */
int sla_last_upload(const sla_ctx *ctx, uint64_t *bytes, uint32_t *value_bytes);

/* Width (2 or 8 bytes) with which the uniform-degree bid scans read each value of the resident CSR.  After a narrow
 * upload at 2 bytes per value the u16 copy stays in HBM beside the widened f64 array, and the grid-wide scans of a CSR
 * whose rows all have the same multiple-of-8 degree read it: 6 instead of 12 bytes per arc, the same doubles after the
 * exact conversion, so the same choices as the reference's f64 scan (ksparse.rs:199-214, symmetric.rs:361-376) bit
 * for bit.  Every other kernel reads the f64 array.  Option "narrow_scan" = 0 turns this off. */
int sla_scan_value_bytes(const sla_ctx *ctx, uint32_t *value_bytes);

/* The narrowing step of that upload on its own (host only, no device): tries to narrow values[0..n) to `tier` bytes
 * each (2: u16, 4: f32) into `out`.  Returns 1 when every value survives the round trip bit for bit (then, with
 * `negate`, `values` has also been negated in place, solver.rs:214-216), 0 when some value does not (`values` is left
 * unchanged), -1 on bad arguments. */
int sla_host_narrow(double *values, size_t n, int tier, void *out, int negate);

/* Same, source arrays already in device memory (device-side generators, multi-GPU shards). */
int sla_upload_csr_device(sla_ctx *ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t *d_row_ptr,
                          const uint32_t *d_column_indices, const double *d_values, uint64_t nnz);

/* Synthetic k-regular instance generated directly in HBM (shapes of the reference's
 * benches/benchmark.rs:49-79 and 16-47): every row has k distinct sorted columns, integer costs uniform in
 * [value_lo, value_hi); planted != 0 adds one arc of a fixed permutation per row so that a perfect matching
 * exists.  sla_generate_host writes the bit-identical instance into host arrays (row_ptr: num_rows+1,
 * cols/vals: num_rows*k). */
int sla_generate_device(sla_ctx *ctx, uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed,
                        uint32_t value_lo, uint32_t value_hi, int planted);
/* Rows [row_begin, row_begin + row_count) of the same global instance, for one rank of a row-partitioned solve. */
int sla_generate_device_shard(sla_ctx *ctx, uint32_t global_rows, uint32_t num_cols, uint32_t k, uint64_t seed,
                              uint32_t value_lo, uint32_t value_hi, int planted, uint32_t row_begin,
                              uint32_t row_count);
int sla_generate_host(uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo,
                      uint32_t value_hi, int planted, uint32_t *row_ptr, uint32_t *column_indices, double *values);
/* Same generator, more knobs: value_dist 0 = uniform integers in [value_lo, value_hi), 1 = floor((hi - lo) * Beta(3,3) + lo)
 * -- the values of the reference's asymmetric bench, floor(700 * Beta(3,3) + 300) (benches/benchmark.rs:60,73); rows
 * [row_begin, row_begin + row_count) of the global instance with row_ptr local to that block (one rank's shard);
 * `threads` host threads.  Both host generators are also exported by libsla_host.so, which is built with g++ and maps no
 * CUDA library (CPU-only processes: the reference arm of bench.py, oracle tests). */
int sla_generate_host_ex(uint32_t global_rows, uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo,
                         uint32_t value_hi, int planted, int value_dist, uint32_t row_begin, uint32_t row_count, int threads,
                         uint32_t *row_ptr, uint32_t *column_indices, double *values);

/* ---- solves.  eps / start_eps: NaN means None; max_iterations: 0 means None (100000, symmetric.rs:190).
 *      Output pointers are host memory (any may be NULL: the result then stays resident and can be fetched
 *      with sla_download_solution): person_to_object[num_rows], object_to_person[num_cols], prices[num_cols]. */
int sla_khosla_solve(sla_ctx *ctx, int maximize, double eps, uint32_t *person_to_object,
                     uint32_t *object_to_person, double *prices, sla_stats *stats);
int sla_forward_solve(sla_ctx *ctx, int maximize, double eps, double start_eps, uint32_t max_iterations,
                      uint32_t *person_to_object, uint32_t *object_to_person, double *prices, sla_stats *stats);
int sla_download_solution(sla_ctx *ctx, uint32_t *person_to_object, uint32_t *object_to_person, double *prices);

/* ---- post-processing on the resident solution (the steps right after the path) ----
 * get_objective (solver.rs:110-142): exact for integer-valued weights; for other weights the device sum is
 * order-independent but not the reference's left-to-right order (the host wrappers keep a sequential sum).
 * ecs_satisfied (solver.rs:154-189).  validate_matching: person_to_object / object_to_person mutually
 * consistent, and the recount of unassigned persons. */
int sla_get_objective(sla_ctx *ctx, double *objective);
int sla_ecs_satisfied(sla_ctx *ctx, double eps, double toleration, int *satisfied);
int sla_validate_matching(sla_ctx *ctx, uint32_t *num_unassigned, int *consistent);

int sla_get_round_profile(sla_ctx *ctx, sla_round_profile *out, size_t capacity, size_t *count);

/* ---- batch of independent instances (BASELINE.json config 4; the reference only clones solvers in a loop,
 *      benches/benchmark.rs:109,137).  Instance b owns rows [row_off[b], row_off[b+1]) and columns
 *      [col_off[b], col_off[b+1]) of the concatenated arrays; column indices are local to the instance;
 *      row_ptr holds global arc offsets (total_rows + 1 entries).  One CTA runs one whole solve. ---- */
int sla_batch_upload(sla_ctx *ctx, uint32_t num_instances, const uint32_t *row_off, const uint32_t *col_off,
                     const uint32_t *row_ptr, const uint32_t *column_indices, const double *values);
int sla_batch_generate_device(sla_ctx *ctx, uint32_t num_instances, uint32_t first_instance_id, uint32_t num_rows,
                              uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo, uint32_t value_hi,
                              int planted);
int sla_batch_solve(sla_ctx *ctx, int algo, int maximize, double eps, double start_eps, uint32_t max_iterations,
                    uint32_t *person_to_object, uint32_t *object_to_person, double *prices,
                    sla_stats *per_instance_stats /* num_instances entries or NULL */, sla_stats *total);

/* ---- row-partitioned single KhoslaSolver instance (BASELINE.json config 5): this context holds persons
 *      [row_begin, row_begin + num_rows) of a global instance with `global_rows` persons (upload the shard with
 *      sla_upload_csr / sla_generate_device first); object state (prices, owners, packed best-bid words) is
 *      replicated on every rank.  One synchronous round is
 *          sla_part_bid    -> all-reduce(MAX) of the best-bid words (uint64, bit 63 clear: valid as int64)
 *          sla_part_claim  -> all-reduce(MAX) of the price candidates (f64, -inf = no bid)
 *          sla_part_assign -> applies all winners to this rank's replica; returns this rank's next queue length.
 *      The caller (sparse_linear_assignment_b200.distributed) owns the collectives and stops when the queue lengths
 *      sum to zero.  global_w_min / global_w_max / global_first_value: min / max over all ranks of
 *      sla_part_local_value_range, and rank 0's first value (sign normalisation, solver.rs:207-216). ---- */
int sla_part_local_value_range(sla_ctx *ctx, double *w_min, double *w_max, double *first_value);
int sla_part_begin(sla_ctx *ctx, int algo, int maximize, uint32_t row_begin, uint32_t global_rows, double eps,
                   double global_w_min, double global_w_max, double global_first_value);
int sla_part_buffers(sla_ctx *ctx, void **d_best_words, void **d_price_candidates, uint64_t *num_words);
int sla_part_bid(sla_ctx *ctx);
int sla_part_claim(sla_ctx *ctx);
int sla_part_assign(sla_ctx *ctx, uint32_t *local_queue_len, uint32_t *local_dropped);
int sla_part_finish(sla_ctx *ctx, uint32_t *person_to_object /* local rows */, uint32_t *object_to_person,
                    double *prices, sla_stats *stats);
/* Sparse exchange (same rounds, same results): instead of all-reducing one word per OBJECT, the ranks all-gather the
 * lists of their local winners (3 x int64 per entry: object, packed word, bits of the exact f64 bid):
 *     sla_part_bid -> sla_part_collect (local winners -> send list, local losers re-queue; returns the count)
 *     all_gather(counts) ; all_gather(send[: 3 * max_count]) into recv        [caller]
 *     sla_part_apply_sparse (global maximum per object over all lists, winners applied to the replica). */
int sla_part_sparse_buffers(sla_ctx *ctx, int world, void **d_send, void **d_recv, void **d_counts, uint64_t *send_capacity);
int sla_part_collect(sla_ctx *ctx, uint32_t *local_winners);
int sla_part_apply_sparse(sla_ctx *ctx, int world, uint32_t max_count, uint32_t *local_queue_len, uint32_t *local_dropped);

/* ---- mesh: one KhoslaSolver instance over the GPUs of one NVLink / NVSwitch domain (BASELINE.json config 5; replaces
 *      the body of KhoslaSolver::solve, ksparse.rs:153-251, for an instance whose rows live on several GPUs).
 *      Persons are row-partitioned (this context holds rows [row_begins[rank], row_begins[rank + 1]), uploaded /
 *      generated as a shard with the GLOBAL number of columns); objects are owner-partitioned in ranges of 2^shift ids
 *      over the same ranks.  All exchange happens inside the round kernels through peer mappings of the ranks' "mesh
 *      blocks": bids are pushed into the owner's HBM, prices are gathered from it, replies and evictions are pushed
 *      back, and the ranks meet in flag barriers written over NVLink -- no collective and no host synchronisation per
 *      round (csrc/sla_mesh.cuh).  Results equal the one-GPU solve bit for bit.
 *          sla_mesh_create   layout + this rank's block (device memory, one cudaMalloc)
 *          sla_ipc_export / sla_ipc_import / sla_ipc_release   the block as a 64-byte handle for the other processes
 *          sla_mesh_connect  every rank's block as addressable from this device (own block, IPC map, or in-process peer)
 *          sla_mesh_begin -> sla_mesh_solve -> sla_mesh_finish   one solve; global_* as for sla_part_begin
 *          sla_mesh_phase / sla_mesh_poll   one kernel of a round at a time, for ranks driven in lockstep by one thread
 *      world <= 8.  A barrier that waits longer than SLA_MESH_TIMEOUT_S (default 20 s) gives up: the solve fails with
 *      SLA_ERR_STATE instead of hanging. ---- */
int sla_mesh_create(sla_ctx *ctx, int rank, int world, const uint32_t *row_begins /* world + 1 */, uint32_t global_cols,
                    void **block, size_t *block_bytes);
int sla_ipc_export(void *dev_ptr, unsigned char *handle64);
int sla_ipc_import(int device, const unsigned char *handle64, void **dev_ptr);
int sla_ipc_release(int device, void *dev_ptr);
int sla_mesh_connect(sla_ctx *ctx, void *const *peer_blocks /* world */, const int *peer_devices /* world or NULL */);
int sla_mesh_begin(sla_ctx *ctx, int maximize, double eps, double global_w_min, double global_w_max,
                   double global_first_value);
int sla_mesh_solve(sla_ctx *ctx);
int sla_mesh_phase(sla_ctx *ctx, int which /* 0 bid, 1 max, 2 resolve, 3 finish */);
int sla_mesh_poll(sla_ctx *ctx, int *done, uint32_t *round, uint32_t *local_queue_len);
/* person_to_object: this rank's rows (global object ids); object_to_person / prices: the objects this rank owns
 * (sla_mesh_owned: first_object .. first_object + num_owned).  stats: this rank's share (the caller adds them up). */
int sla_mesh_finish(sla_ctx *ctx, uint32_t *person_to_object, uint32_t *object_to_person, double *prices, sla_stats *stats);
int sla_mesh_owned(sla_ctx *ctx, uint32_t *shard_objects, uint32_t *num_owned, uint32_t *first_object, uint32_t *first_row);
/* Duration of the first round's bid kernel of the last solve (event-record nodes inside the first graph), and this rank's
 * share of get_objective (solver.rs:110-142; exact for integer weights) -- the caller adds the shares up. */
int sla_mesh_round1_ms(sla_ctx *ctx, float *bid_ms);
/* Development aid (environment SLA_MESH_TIMELINE=1 at sla_mesh_create): %globaltimer stamps of the last solve, 8 per round. */
int sla_mesh_timeline(sla_ctx *ctx, unsigned long long *out, size_t capacity);
int sla_mesh_objective(sla_ctx *ctx, double *objective);

#ifdef __cplusplus
}
#endif
#endif /* SLA_B200_H */
