// sla_api.cu -- host side of libsla_b200.so: context, CSR mirror, solve drivers (CUDA-graph super-rounds or a
// host-driven loop), post-processing, and the extern "C" boundary declared in include/sla.h.
#include <algorithm>
#include <cmath>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#include <cstdio>
#include <cstdlib>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only NVTX v3: ranges cost a few ns unless a profiler is attached

#include "../../include/sla.h"
#include "sla_kernels.cuh"
#include "sla_host_impl.h"

using namespace sla;

// ---------------------------------------------------------------------------------------------------------
namespace {

thread_local std::string g_create_error;

// NVTX ranges around the phases of the path -- the device-side counterpart of the reference's `trace!` hooks
// (ksparse.rs:182,189-190, symmetric.rs:406-407): upload, solve, the graph launches inside it, download.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

struct GraphSlot {
    cudaGraphExec_t exec = nullptr;
    uint64_t generation = 0;
    int lpr = 0;
    int super_rounds = 0;
    uint32_t regular_k = 0;
    bool smem_prices = false;
    bool tail_only = false;
    uint32_t tail_smem_bytes = 0;
    uint32_t tail_max = 0;
    bool khosla_phases = false;
    bool zero_first = false;
    size_t l2_bytes = 0;
    const void* vals16 = nullptr;   // u16 value mirror the regular bid kernels were captured with (nullptr: f64 values)
    uint32_t learned_wide = 0;      // > 0: first graph of a plain Khosla solve shaped as `learned_wide` wide rounds + one tail launch
    uint32_t own_max = 0;           // DevState::tail_own_max the graph was captured for; the cluster engine is in it when own_max < tail_max
};

}  // namespace

struct sla_batch_state;
struct sla_part_state;
struct sla_mesh_state;

// Persistent host workers of one context for the in-place negation of the caller's `values` (solver.rs:214-216):
// chunk w of the array is negated by worker w as soon as the H2D copy of that chunk has completed (one event per chunk),
// so the negation trails the upload by one chunk instead of starting after it.  Created lazily; the workers sleep on a
// condition variable between jobs.
struct NegPool {
    static constexpr int kMax = 16;
    static constexpr int kMaxPieces = 4096;
    enum Mode : int { NEGATE_AFTER_COPY = 0, NARROW = 1 };
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    uint64_t job_id = 0;
    int pending = 0;
    bool stop = false;
    int mode = NEGATE_AFTER_COPY;
    double* values = nullptr;
    size_t chunk = 0, total = 0;
    int nchunks = 0;
    int device = 0;
    cudaEvent_t ev[kMax] = {};
    // NARROW: lossless narrowing of `values` into `stage` (tier 2: u16, tier 4: f32), piece by piece in array order;
    // done[p] is set (release) when piece p is staged, `narrow_fail` when a value does not survive the round trip
    void* stage = nullptr;
    int tier = 0;
    size_t piece = 0;
    int npieces = 0;
    bool narrow_negate = false;  // NARROW: negate each piece in place in the very pass that stages it
    int narrow_workers = kMax;   // workers 0 .. narrow_workers-1 take pieces (co-located ranks share the cores)
    std::atomic<int> next_piece{0};
    std::atomic<int> narrow_fail{0};
    std::atomic<unsigned char> done[kMaxPieces];

    // true when every value of [lo, hi) is an integer in [0, 65535] with a clear sign bit (so that widening restores the
    // exact bit pattern); out = the values as u16 (streaming stores: the staging block is only read by the DMA engine).
    // With `negate` the values are negated in place in the same pass (and restored if the piece turns out not to be
    // representable).
    static bool narrow_u16(double* v, uint16_t* out, size_t lo, size_t hi, bool negate) {
        size_t i = lo;
        bool good = true;
#if defined(__SSE2__)
        if ((((uintptr_t)(out + lo)) & 15u) == 0) {
            __m128i bad = _mm_setzero_si128();
            __m128d ok = _mm_castsi128_pd(_mm_set1_epi32(-1)), sgn = _mm_setzero_pd();
            const __m128d sign = _mm_set1_pd(-0.0);
            for (; i + 8 <= hi; i += 8) {
                const __m128d a = _mm_loadu_pd(v + i), b = _mm_loadu_pd(v + i + 2), c = _mm_loadu_pd(v + i + 4), d = _mm_loadu_pd(v + i + 6);
                const __m128i ia = _mm_cvttpd_epi32(a), ib = _mm_cvttpd_epi32(b), ic = _mm_cvttpd_epi32(c), id = _mm_cvttpd_epi32(d);
                ok = _mm_and_pd(ok, _mm_and_pd(_mm_and_pd(_mm_cmpeq_pd(_mm_cvtepi32_pd(ia), a), _mm_cmpeq_pd(_mm_cvtepi32_pd(ib), b)),
                                               _mm_and_pd(_mm_cmpeq_pd(_mm_cvtepi32_pd(ic), c), _mm_cmpeq_pd(_mm_cvtepi32_pd(id), d))));
                sgn = _mm_or_pd(sgn, _mm_or_pd(_mm_or_pd(a, b), _mm_or_pd(c, d)));
                const __m128i lo4 = _mm_unpacklo_epi64(ia, ib), hi4 = _mm_unpacklo_epi64(ic, id);   // 4 + 4 int32
                bad = _mm_or_si128(bad, _mm_or_si128(lo4, hi4));
                const __m128i p = _mm_packs_epi32(_mm_srai_epi32(_mm_slli_epi32(lo4, 16), 16), _mm_srai_epi32(_mm_slli_epi32(hi4, 16), 16));
                _mm_stream_si128((__m128i*)(out + i), p);
                if (negate) {
                    _mm_storeu_pd(v + i, _mm_xor_pd(a, sign)); _mm_storeu_pd(v + i + 2, _mm_xor_pd(b, sign));
                    _mm_storeu_pd(v + i + 4, _mm_xor_pd(c, sign)); _mm_storeu_pd(v + i + 6, _mm_xor_pd(d, sign));
                }
            }
            _mm_sfence();
            const bool range_ok = _mm_movemask_epi8(_mm_cmpeq_epi32(_mm_and_si128(bad, _mm_set1_epi32((int)0xFFFF0000u)), _mm_setzero_si128())) == 0xFFFF;
            good = range_ok && _mm_movemask_pd(ok) == 3 && _mm_movemask_pd(sgn) == 0;
        }
#endif
        for (; good && i < hi; ++i) {
            const double x = v[i];
            if (!(x >= 0.0 && x <= 65535.0)) { good = false; break; }
            const uint16_t q = (uint16_t)x;
            const double back = (double)q;
            if (__builtin_memcmp(&back, &x, 8) != 0) { good = false; break; }
            out[i] = q;
            if (negate) v[i] = -x;
        }
        if (!good && negate) for (size_t j = lo; j < i; ++j) v[j] = -v[j];   // put the piece back as it was
        return good;
    }
    // same for f32: bit-exact round trip
    static bool narrow_f32(double* v, float* out, size_t lo, size_t hi, bool negate) {
        size_t i = lo;
        bool good = true;
#if defined(__SSE2__)
        if ((((uintptr_t)(out + lo)) & 15u) == 0) {
            __m128i same = _mm_set1_epi32(-1);
            const __m128d sign = _mm_set1_pd(-0.0);
            for (; i + 4 <= hi; i += 4) {
                const __m128d a = _mm_loadu_pd(v + i), b = _mm_loadu_pd(v + i + 2);
                const __m128 fa = _mm_cvtpd_ps(a), fb = _mm_cvtpd_ps(b);
                same = _mm_and_si128(same, _mm_and_si128(_mm_cmpeq_epi32(_mm_castpd_si128(_mm_cvtps_pd(fa)), _mm_castpd_si128(a)),
                                                         _mm_cmpeq_epi32(_mm_castpd_si128(_mm_cvtps_pd(fb)), _mm_castpd_si128(b))));
                _mm_stream_ps(out + i, _mm_movelh_ps(fa, fb));
                if (negate) { _mm_storeu_pd(v + i, _mm_xor_pd(a, sign)); _mm_storeu_pd(v + i + 2, _mm_xor_pd(b, sign)); }
            }
            _mm_sfence();
            good = _mm_movemask_epi8(same) == 0xFFFF;
        }
#endif
        for (; good && i < hi; ++i) {
            const double x = v[i];
            const float q = (float)x;
            const double back = (double)q;
            if (__builtin_memcmp(&back, &x, 8) != 0) { good = false; break; }
            out[i] = q;
            if (negate) v[i] = -x;
        }
        if (!good && negate) for (size_t j = lo; j < i; ++j) v[j] = -v[j];   // put the piece back as it was
        return good;
    }

    void worker(int w) {
        cudaSetDevice(device);
        uint64_t seen = 0;
        while (true) {
            double* v; size_t lo, hi; bool mine; int md;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || job_id != seen; });
                if (stop) return;
                seen = job_id;
                md = mode;
                mine = md == NARROW ? w < narrow_workers : w < nchunks;
                v = values;
                lo = (size_t)w * chunk;
                hi = lo + chunk < total ? lo + chunk : total;
            }
            if (!mine) continue;
            if (md == NARROW) {
                while (!narrow_fail.load(std::memory_order_relaxed)) {
                    const int pc = next_piece.fetch_add(1, std::memory_order_relaxed);
                    if (pc >= npieces) break;
                    const size_t plo = (size_t)pc * piece, phi = plo + piece < total ? plo + piece : total;
                    const bool ok = tier == 2 ? narrow_u16(v, (uint16_t*)stage, plo, phi, narrow_negate)
                                              : narrow_f32(v, (float*)stage, plo, phi, narrow_negate);
                    if (!ok) { narrow_fail.store(1, std::memory_order_release); break; }
                    done[pc].store(1, std::memory_order_release);
                }
            } else {
                cudaEventSynchronize(ev[w]);               // the DMA engine has finished reading this chunk
                for (size_t i = lo; i < hi; ++i) v[i] = -v[i];
            }
            {
                std::lock_guard<std::mutex> lk(m);
                if (--pending == 0) cv_done.notify_all();
            }
        }
    }
    bool start(int dev) {
        if (!threads.empty()) return true;
        device = dev;
        for (int i = 0; i < kMax; ++i)
            if (cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) return false;
        unsigned hw = std::thread::hardware_concurrency();
        int n = hw ? (int)hw : 4;
        if (n > kMax) n = kMax;
        for (int w = 0; w < n; ++w) threads.emplace_back([this, w] { worker(w); });
        return true;
    }
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
    void shutdown() {
        if (threads.empty()) return;
        wait();
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_job.notify_all();
        for (auto& t : threads) t.join();
        threads.clear();
        for (auto& e : ev) if (e) { cudaEventDestroy(e); e = nullptr; }
    }
};

struct sla_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string err;

    // capacities (elements)
    size_t cap_rows = 0, cap_cols = 0, cap_arcs = 0;
    // CSR mirror
    uint32_t* d_row_ptr = nullptr;
    uint32_t* d_cols = nullptr;
    double* d_vals = nullptr;
    // solve state
    double* d_prices = nullptr;
    uint32_t* d_p2o = nullptr;
    uint32_t* d_o2p = nullptr;
    unsigned long long* d_best = nullptr;
    uint32_t* d_queue[2] = {nullptr, nullptr};
    uint32_t* d_slot_obj = nullptr;
    double* d_slot_bid = nullptr;
    DevState* d_state = nullptr;
    DevCsrStats* d_csr_stats = nullptr;
    uint32_t* d_scratch = nullptr;   // 16 words
    double* d_partial = nullptr;     // objective partial sums, one per block
    // pinned staging
    unsigned char* h_small = nullptr;     // small instances: one block for the CSR on its way up and the results on their way down
    size_t h_small_cap = 0;
    cudaEvent_t ev_small = nullptr;       // the last H2D copy out of h_small has completed
    cudaStream_t stream_ride = nullptr;   // u16 uploads: the kernels that ride behind the pieces run here, so that the copies
    cudaEvent_t ev_ride = nullptr;        // on `stream` stay back to back (created on first use)
    bool small_pending = false;
    // large uploads: values that survive a round trip through u16 / f32 cross PCIe narrow and are widened on the device
    void* h_narrow = nullptr;             // page-locked staging (tier bytes per value)
    size_t h_narrow_cap = 0;
    void* d_narrow = nullptr;
    size_t d_narrow_cap = 0;
    int opt_narrow_upload = 1;
    int opt_narrow_scan = 1;       // regular bid kernels read the u16 mirror a narrow upload left in HBM (6 B per arc)
    bool vals16_valid = false;     // d_narrow holds the u16 image of every value in d_vals
    uint64_t last_upload_bytes = 0;       // bytes the last upload moved host -> device
    uint32_t last_upload_value_bytes = 8; // 2 / 4 / 8: width the values crossed PCIe with
    unsigned char* h_dl = nullptr;        // small instances: results land here behind each graph launch (one sync per launch)
    size_t h_dl_cap = 0;
    DevState* h_state = nullptr;
    DevCsrStats* h_csr_stats = nullptr;
    uint32_t* h_scratch = nullptr;
    double* h_partial = nullptr;

    // problem
    bool has_csr = false, has_solution = false, best_dirty = true;
    bool l2_attr_set = false;  // the stream carries a persisting access-policy window (apply_l2_policy)
    uint32_t n_rows = 0, n_cols = 0;
    uint64_t nnz = 0;
    double v_min = 0, v_max = 0, first_value = 0;
    int dev_sign = 1;   // effective values = dev_sign * uploaded values (in-place negation of solver.rs:214-216)
    int lpr = 4;
    uint32_t regular_k = 0;   // all rows have this many arcs (multiple of 8): the regular bid kernel is used
    int lpr8 = 1;
    int opt_regular = 1;
    int opt_small_path = 1;    // small instances: host-side statistics on the way up, results behind the graph on the way down
    int opt_wide_first = 1;    // first round of a small plain-Khosla solve on the wide kernels (solve_tail_max)
    int opt_l2_persist = 1;    // access-policy window (persisting) over the bid words when they fit the carve-out
    bool l2_active = false;    // this context holds a share of the device's persisting-L2 carve-out
    size_t l2_bytes = 0;       // window size the stream attribute / captured graphs were set up with
    size_t l2_persist_max = 0, l2_window_max = 0;
    int opt_stream_scan = 0;   // first-round scan through the TMA pipeline (bid_stream_kernel): opt-in, measured 4 % slower
    int opt_smem_prices = 1, opt_smem_owners = 1, opt_khosla_scaling = 1;
    int opt_prune = 1;         // bound-pruned gather in the uniform-degree scans of rounds with >= 32 Ki bidders
    int opt_upload_ride = 1;   // u16 uploads: statistics + widening behind every run of pieces instead of two passes at the end
    int opt_cluster_engine = 0; // opt-in ("cluster_engine" = 1; large M, no shared-memory price mirror): queues of kTailSlots+1 ..
                                // kMidMax bidders run in mid_kernel.  Bit-identical, but measured no faster than the grid-wide pair +
                                // single-CTA engine it replaces (cfg3 0.181 vs 0.172 ms, cfg5 2.73 vs 2.71 ms: a round is ~7 us of
                                // dependent memory round trips either way; profiles/r02_cluster_engine_probe.jsonl), hence off
    int opt_cluster_handover = (int)kTailSlots;   // ... which hands over to the single-CTA engine at this many bidders
    int opt_learn_shape = 1;   // plain Khosla: the first graph of a solve has as many wide rounds as the previous solve of the
                               // same resident CSR needed, followed by one tail launch (no no-op launches in between)
    int opt_mesh_tail = 1;     // mesh engine: the persistent one-block tail engine runs the short rounds
    int opt_prezero_best = 0;  // development: zero the bid words in front of every solve (the RED targets then sit in the L2)
    uint32_t learned_wide = 0; // wide rounds of the last plain Khosla solve of the resident CSR (0: unknown)
    // tail-engine plan of the current instance (plan_tail): what is mirrored in shared memory and how many bidders fit
    bool tail_smem_prices = false;   // object prices mirrored
    uint32_t tail_own_mode = 0;      // owners mirrored: 0 no, 1 u32, 2 u16
    uint32_t tail_cap = 1024;        // capacity of the queue arrays (power of two)
    bool super_rounds_user = false;  // "super_rounds" was set through sla_set_option
    bool tail_max_user = false;      // "tail_max" was set through sla_set_option
    uint32_t tail_max_eff = 1024;    // bidders at or below which the tail engine runs (min(option tail_max, tail_cap))
    uint32_t tail_smem_bytes = 0;    // dynamic shared memory of a tail launch

    // options
    int opt_graph = 1, opt_tail_max = 1024, opt_skip_zero = 1, opt_profile = 0, opt_super_rounds = 6, opt_profile_repeat = 1;
    double opt_timeout_s = 900.0;   // wall-clock guard of one solve ("timeout_s" option / SLA_TIMEOUT_S)
    uint64_t generation = 1;   // bumped whenever a device buffer is reallocated
    GraphSlot graphs[4];   // [forward * 2 + first-launch-of-a-solve]
    int grid_wide = 0;

    std::vector<sla_round_profile> profile;
    int opt_profile_graph = 1;             // "profile": bracket each wide scan inside one small graph launch (see solve_common)
    cudaGraphExec_t profile_exec = nullptr;
    bool profile_graph_failed = false;
    NegPool neg;                           // host workers negating the caller's `values` (sla_upload_csr_negating)

    sla_batch_state* batch = nullptr;
    sla_part_state* part = nullptr;
    sla_mesh_state* mesh = nullptr;
};

namespace {

void join_workers(sla_ctx* c) { c->neg.wait(); }

struct WorkerJoinGuard {   // whatever path leaves a solve, the caller's host arrays are quiescent afterwards
    sla_ctx* c;
    ~WorkerJoinGuard() { if (c) join_workers(c); }
};

int fail(sla_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CU(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e__ = (call);                                                                        \
        if (e__ != cudaSuccess)                                                                          \
            return fail(ctx, SLA_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__));         \
    } while (0)

template <class T>
int dev_alloc(sla_ctx* ctx, T** p, size_t n) {
    if (*p) { cudaFree(*p); *p = nullptr; }
    cudaError_t e = cudaMalloc((void**)p, (n ? n : 1) * sizeof(T));
    if (e != cudaSuccess) {
        *p = nullptr;
        return fail(ctx, SLA_ERR_ALLOC, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    return SLA_OK;
}

Params make_params(const sla_ctx* c) {
    Params p;
    p.row_ptr = c->d_row_ptr; p.cols = c->d_cols; p.vals = c->d_vals;
    p.prices = c->d_prices; p.p2o = c->d_p2o; p.o2p = c->d_o2p; p.best = c->d_best;
    p.queue[0] = c->d_queue[0]; p.queue[1] = c->d_queue[1];
    p.slot_obj = c->d_slot_obj; p.slot_bid = c->d_slot_bid; p.st = c->d_state;
    p.vals16 = c->vals16_valid ? static_cast<const uint16_t*>(c->d_narrow) : nullptr;
    return p;
}

void drop_graphs(sla_ctx* c) {
    for (auto& g : c->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = GraphSlot();
    }
}

int ensure_capacity(sla_ctx* ctx, size_t rows, size_t cols, size_t arcs) {
    bool changed = false;
    if (rows > ctx->cap_rows || !ctx->d_row_ptr) {
        size_t n = rows > ctx->cap_rows ? rows : ctx->cap_rows;
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->d_row_ptr, n + 8))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_p2o, n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_queue[0], n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_queue[1], n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_slot_obj, n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_slot_bid, n))) return rc;
        ctx->cap_rows = n;
        changed = true;
    }
    if (cols > ctx->cap_cols || !ctx->d_prices) {
        size_t n = cols > ctx->cap_cols ? cols : ctx->cap_cols;
        int rc;
        if ((rc = dev_alloc(ctx, &ctx->d_prices, n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_o2p, n))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_best, n))) return rc;
        ctx->cap_cols = n;
        ctx->best_dirty = true;
        changed = true;
    }
    if (arcs > ctx->cap_arcs || !ctx->d_cols) {
        size_t n = arcs > ctx->cap_arcs ? arcs : ctx->cap_arcs;
        int rc;
        // +8: scan_row reads whole aligned 4-arc chunks, so the tail of the last row may be over-read
        if ((rc = dev_alloc(ctx, &ctx->d_cols, n + 8))) return rc;
        if ((rc = dev_alloc(ctx, &ctx->d_vals, n + 8))) return rc;
        CU(cudaMemsetAsync(ctx->d_cols, 0, (n + 8) * sizeof(uint32_t), ctx->stream));
        CU(cudaMemsetAsync(ctx->d_vals, 0, (n + 8) * sizeof(double), ctx->stream));
        ctx->cap_arcs = n;
        changed = true;
    }
    if (changed) {
        ctx->generation += 1;
        drop_graphs(ctx);
        ctx->has_csr = false;
        ctx->has_solution = false;
    }
    return SLA_OK;
}

int pick_lpr(uint64_t nnz, uint32_t n_rows) {
    const double avg = n_rows ? (double)nnz / (double)n_rows : 1.0;
    int lpr = 1;
    while (lpr < 32 && (double)(lpr * 4) < avg) lpr *= 2;
    return lpr;
}

// ---- kernel dispatch on lanes-per-row -----------------------------------------------------------------
#ifndef SLA_REG_OCC_ZERO
#define SLA_REG_OCC_ZERO 3
#endif
// First round of a solve on a regular CSR with K <= 256: the TMA pipeline (rows of a tile are one contiguous range).
bool use_stream_scan(const sla_ctx* c) { return c->opt_stream_scan && c->regular_k != 0 && c->regular_k <= 256u; }
// The values crossed PCIe as u16 and that copy is still in HBM: the uniform-degree scans read it instead of the
// widened f64 array (same doubles after conversion, half the bytes per arc).
const void* narrow_scan_ptr(const sla_ctx* c) {
    return (c->opt_narrow_scan && c->vals16_valid && c->regular_k != 0 && c->regular_k <= 32768u && c->opt_regular) ? c->d_narrow : nullptr;
}

template <int MODE>
void launch_bid_regular_m(sla_ctx* c, const Params& p) {
    if (MODE == PRICE_ZERO && use_stream_scan(c)) {
        const int grid = c->num_sms * 2;
        switch (c->lpr8) {
            case 1: bid_stream_kernel<1><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
            case 2: bid_stream_kernel<2><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
            case 4: bid_stream_kernel<4><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
            case 8: bid_stream_kernel<8><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
            case 16: bid_stream_kernel<16><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
            default: bid_stream_kernel<32><<<grid, kStreamThreads + 32, kStreamSmemBytes, c->stream>>>(p); break;
        }
        return;
    }
    const int grid_wide = (MODE == PRICE_ZERO) ? c->num_sms * SLA_REG_OCC_ZERO : c->grid_wide;
    if (narrow_scan_ptr(c) && p.vals16) {
        const int grid_wide = (MODE == PRICE_ZERO) ? c->num_sms * SLA_KEY_OCC : c->grid_wide;
        switch (c->lpr8) {
            case 1: bid_regular_kernel<1, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
            case 2: bid_regular_kernel<2, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
            case 4: bid_regular_kernel<4, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
            case 8: bid_regular_kernel<8, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
            case 16: bid_regular_kernel<16, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
            default: bid_regular_kernel<32, MODE, true><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        }
        return;
    }
    switch (c->lpr8) {
        case 1: bid_regular_kernel<1, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        case 2: bid_regular_kernel<2, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        case 4: bid_regular_kernel<4, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        case 8: bid_regular_kernel<8, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        case 16: bid_regular_kernel<16, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        default: bid_regular_kernel<32, MODE><<<grid_wide, kWideThreads, 0, c->stream>>>(p); break;
    }
}

bool use_regular(const sla_ctx* c) { return c->regular_k != 0 && c->opt_regular; }

// which: 0 bid, 1 assign, 2 tail, 3 ecs, 4 phase_apply, 5 cluster engine.  zero_first: this bid launch is the first round of a
// solve and all prices are known to be exactly 0 (only meaningful for the regular bid kernel; the generic one
// reads the flag from the device state).
template <int LPR>
void launch_one_t(sla_ctx* c, const Params& p, int which, bool zero_first) {
    switch (which) {
        case 0:
            if (use_regular(c)) {
                if (zero_first) launch_bid_regular_m<PRICE_ZERO>(c, p);
                else launch_bid_regular_m<PRICE_LDG>(c, p);
            } else {
                bid_wide_kernel<LPR><<<c->grid_wide, kWideThreads, 0, c->stream>>>(p);
            }
            break;
        case 1: assign_wide_kernel<<<c->grid_wide, kWideThreads, 0, c->stream>>>(p, zero_first ? 1 : 0); break;
        case 2:
            if (c->tail_smem_prices) tail_kernel<LPR, true><<<1, kTailThreads, c->tail_smem_bytes, c->stream>>>(p);
            else tail_kernel<LPR, false><<<1, kTailThreads, c->tail_smem_bytes, c->stream>>>(p);
            break;
        case 3: ecs_kernel<LPR><<<c->grid_wide, kWideThreads, 0, c->stream>>>(p); break;
        case 5: mid_kernel<LPR><<<kMidCtas, kMidThreads, 0, c->stream>>>(p); break;   // one cluster (__cluster_dims__)
        default: phase_apply_kernel<<<c->grid_wide, kWideThreads, 0, c->stream>>>(p); break;
    }
}
void launch_one(sla_ctx* c, const Params& p, int which, bool zero_first = false) {
    switch (c->lpr) {
        case 1: launch_one_t<1>(c, p, which, zero_first); break;
        case 2: launch_one_t<2>(c, p, which, zero_first); break;
        case 4: launch_one_t<4>(c, p, which, zero_first); break;
        case 8: launch_one_t<8>(c, p, which, zero_first); break;
        case 16: launch_one_t<16>(c, p, which, zero_first); break;
        default: launch_one_t<32>(c, p, which, zero_first); break;
    }
}

// tail_only: the instance has at most tail_max persons, so the queue can never be long enough for the wide pair --
// leave those two (no-op) launches out of the super-round.
// Khosla rounds run under an eps-schedule on square instances (finish_if_possible): the phase kernel is then part of
// the Khosla super-round as well.
bool khosla_phases(const sla_ctx* c) { return c->opt_khosla_scaling && c->n_rows == c->n_cols; }

// The cluster engine takes the queues between the single-CTA engine's reach and kMidMax when the tail engine has no
// shared-memory price mirror (large M); an explicit "tail_max" option keeps the two-engine split it asks for.
bool use_mid(const sla_ctx* c) {
    return c->opt_cluster_engine && c->has_csr && !c->tail_smem_prices && !c->tail_max_user && c->tail_max_eff >= kTailSlots;
}

void launch_super_round(sla_ctx* c, const Params& p, bool forward, bool zero_first, bool tail_only) {
    if (!tail_only) {
        launch_one(c, p, 0, zero_first);
        // first round with all prices zero: prices / owners / assignment are initialised behind the scan
        if (zero_first) first_assign_objects_kernel<<<c->grid_wide, kWideThreads, 0, c->stream>>>(p);
        launch_one(c, p, 1, zero_first);
    }
    if (use_mid(c)) launch_one(c, p, 5);
    launch_one(c, p, 2);
    if (forward) {
        launch_one(c, p, 3);
        launch_one(c, p, 4);
    } else if (khosla_phases(c)) {
        launch_one(c, p, 4);
    }
}

int kernels_per_super_round(const sla_ctx* c, bool forward, bool tail_only) {
    return (forward ? 5 : (khosla_phases(c) ? 4 : 3)) - (tail_only ? 2 : 0) + (use_mid(c) ? 1 : 0);
}

// Plain Khosla rounds (no eps-schedule) finish in a handful of rounds whose first one has all N bidders: when N is
// within the tail engine's reach but large (cfg1: 1,000 persons x 32 arcs), that first round is better spread over
// the whole GPU -- the tail engine takes over from round 2 (a few per cent of N).  "wide_first" = 0 turns this off.
constexpr uint32_t kWideFirstMinRows = 640;
uint32_t solve_tail_max(const sla_ctx* c, bool forward) {
    if (use_mid(c)) return kMidMax;   // the grid-wide pair stops here; (solve_own_max, kMidMax] is the cluster engine's
    if (!forward && !khosla_phases(c) && c->opt_wide_first && c->n_rows > kWideFirstMinRows && c->n_rows <= c->tail_max_eff)
        return c->n_rows / 2u;
    return c->tail_max_eff;
}

// DevState::tail_own_max: what the single-CTA engine takes
uint32_t solve_own_max(const sla_ctx* c, bool forward) {
    if (!use_mid(c)) return solve_tail_max(c, forward);
    const uint32_t h = (uint32_t)c->opt_cluster_handover;
    return h < c->tail_max_eff ? h : c->tail_max_eff;
}

bool is_tail_only(const sla_ctx* c, bool forward) { return c->n_rows <= solve_tail_max(c, forward); }

// Khosla on a tail-only instance finishes inside the first tail launch: further super-rounds would be pure no-ops.
// Small rectangular Khosla instances need one wide round before the tail engine finishes them: two super-rounds.
int super_rounds_for(const sla_ctx* c, bool forward) {
    if (!forward && !khosla_phases(c)) {
        if (is_tail_only(c, forward)) return 1;
        // wide first round (solve_tail_max = N / 2): the tail engine finishes the solve behind it unless more than half of
        // the persons lost round 1 -- then the continuation graph takes over
        if (c->n_rows <= c->tail_max_eff) return 1;
        if (c->n_rows <= 4u * c->tail_max_eff && c->opt_super_rounds > 2) return 2;
    }
    // very large instances need more wide rounds before the tail engine can take over (cfg5: 10): a spare super-round
    // costs 3.5 us of control-step launches, a second graph launch a host round trip
    if (!c->super_rounds_user && c->n_rows > (1u << 22) && c->opt_super_rounds < 12) return 12;
    return c->opt_super_rounds;
}

// The bid words (8 B per object) are the RED.MAX target of every bid and are swept once by the first-round object
// pass, while 12 B per arc of CSR stream past them: when they are too large to survive that stream in the L2 by
// themselves but small enough for the set-aside carve-out (8 MB .. persistingL2CacheMaxSize, i.e. 1 M .. 9.9 M objects on
// B200), they are marked persisting (access-policy window on the stream and on the captured kernel nodes).  cfg3: bid
// scan 49 -> 42 us, first-round assignment 36 -> 32 us, solve 0.203 -> 0.195 ms.  The carve-out is a device-wide
// setting: it is raised only as far as needed, and the last context that used it on a device restores the previous
// limit and resets the persisting lines.  Option "l2_persist" = 0 turns it off.
std::mutex g_l2_mutex;
int g_l2_users[64] = {};
size_t g_l2_prev_limit[64] = {};

bool l2_window(const sla_ctx* c, cudaAccessPolicyWindow* w) {
    if (!c->opt_l2_persist || !c->d_best || !c->l2_persist_max || !c->l2_window_max || !c->has_csr) return false;
    const size_t bytes = (size_t)c->n_cols * 8u;
    if (bytes < ((size_t)8 << 20) || bytes > c->l2_persist_max || bytes > c->l2_window_max) return false;
    w->base_ptr = c->d_best;
    w->num_bytes = bytes;
    w->hitRatio = 1.0f;
    w->hitProp = cudaAccessPropertyPersisting;
    w->missProp = cudaAccessPropertyNormal;
    return true;
}

void release_l2_policy(sla_ctx* c) {
    if (!c->l2_active) return;
    c->l2_active = false;
    std::lock_guard<std::mutex> lk(g_l2_mutex);
    const int d = c->device & 63;
    if (--g_l2_users[d] == 0) {
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, g_l2_prev_limit[d]);
        cudaGetLastError();
    }
}

// Called whenever the instance (n_cols), the buffers or the option change.
void apply_l2_policy(sla_ctx* c) {
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof v);
    if (l2_window(c, &v.accessPolicyWindow)) {
        std::lock_guard<std::mutex> lk(g_l2_mutex);
        const int d = c->device & 63;
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        if (!c->l2_active) {
            if (g_l2_users[d]++ == 0) g_l2_prev_limit[d] = cur;
            c->l2_active = true;
        }
        if (cur < v.accessPolicyWindow.num_bytes) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, v.accessPolicyWindow.num_bytes);
        c->l2_bytes = v.accessPolicyWindow.num_bytes;
        c->l2_attr_set = true;
    } else {
        release_l2_policy(c);
        c->l2_bytes = 0;
        if (!c->l2_attr_set) return;     // the stream never had a window: nothing to take back (every small upload comes through here)
        c->l2_attr_set = false;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &v);
    cudaGetLastError();
}

// The persisting-L2 window over the bid words on every kernel node of a captured graph (the stream attribute alone
// is not relied upon for captured work).
void window_on_kernel_nodes(const sla_ctx* ctx, cudaGraph_t graph) {
    cudaKernelNodeAttrValue av;
    memset(&av, 0, sizeof av);
    if (!l2_window(ctx, &av.accessPolicyWindow)) return;
    size_t nn = 0;
    cudaGraphGetNodes(graph, nullptr, &nn);
    std::vector<cudaGraphNode_t> nodes(nn);
    if (nn) cudaGraphGetNodes(graph, nodes.data(), &nn);
    for (cudaGraphNode_t nd : nodes) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nd, &t) == cudaSuccess && t == cudaGraphNodeTypeKernel)
            cudaGraphKernelNodeSetAttribute(nd, cudaKernelNodeAttributeAccessPolicyWindow, &av);
    }
    cudaGetLastError();
}

// Plain Khosla solves (no eps-schedule) of a CSR that stays resident are deterministic: the number of wide rounds the
// previous solve needed before the tail engine took over is the number this one needs.  The first graph of such a solve
// is then captured as exactly that many (bid, assign) pairs followed by ONE tail launch, instead of a tail launch behind
// every pair that returns at once (2.5 us each: cfg3 has five of them).  Every kernel still decides from the device
// state whether it has work, so a wrong guess (another eps, another sign) costs time, never correctness: continuation
// graphs of the default shape finish the solve.
uint32_t learned_shape(const sla_ctx* c, bool forward, bool first_graph) {
    if (!first_graph || forward || khosla_phases(c) || !c->opt_learn_shape || is_tail_only(c, forward)) return 0u;
    return (c->learned_wide >= 1u && c->learned_wide <= 64u) ? c->learned_wide : 0u;
}

int graph_kernel_count(const sla_ctx* c, bool forward, bool first_graph, bool late_init) {
    const uint32_t lw = learned_shape(c, forward, first_graph);
    int n = lw ? (int)(2u * lw + 1u + (use_mid(c) ? 1u : 0u))
               : super_rounds_for(c, forward) * kernels_per_super_round(c, forward, is_tail_only(c, forward));
    if (late_init && first_graph) n += 1;   // first_assign_objects_kernel sits in the first graph
    return n;
}

int get_graph(sla_ctx* ctx, bool forward, bool zero_first, bool first_graph, cudaGraphExec_t* out) {
    GraphSlot& g = ctx->graphs[(forward ? 2 : 0) + (first_graph ? 1 : 0)];
    const uint32_t reg_key = use_regular(ctx) ? ctx->regular_k : 0u;
    const bool tail_only = is_tail_only(ctx, forward);
    const int n_super = super_rounds_for(ctx, forward);
    const uint32_t lw = learned_shape(ctx, forward, first_graph);
    if (g.exec && g.generation == ctx->generation && g.lpr == ctx->lpr && g.super_rounds == n_super && g.learned_wide == lw &&
        g.regular_k == reg_key && g.smem_prices == ctx->tail_smem_prices && g.tail_only == tail_only &&
        g.tail_smem_bytes == ctx->tail_smem_bytes && g.tail_max == solve_tail_max(ctx, forward) && g.khosla_phases == khosla_phases(ctx) &&
        g.l2_bytes == ctx->l2_bytes && g.vals16 == narrow_scan_ptr(ctx) && g.zero_first == zero_first &&
        g.own_max == solve_own_max(ctx, forward)) {
        *out = g.exec;
        return SLA_OK;
    }
    if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
    const Params p = make_params(ctx);
    cudaGraph_t graph = nullptr;
    CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    if (lw) {
        for (uint32_t r = 0; r < lw; ++r) {
            const bool zf = zero_first && r == 0;
            launch_one(ctx, p, 0, zf);
            if (zf) first_assign_objects_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p);
            launch_one(ctx, p, 1, zf);
        }
        if (use_mid(ctx)) launch_one(ctx, p, 5);
        launch_one(ctx, p, 2);
    } else {
        for (int r = 0; r < n_super; ++r) launch_super_round(ctx, p, forward, zero_first && r == 0, tail_only);
    }
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
    if (e != cudaSuccess) return fail(ctx, SLA_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    window_on_kernel_nodes(ctx, graph);
    e = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        g.exec = nullptr;
        return fail(ctx, SLA_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
    }
    g.generation = ctx->generation;
    g.lpr = ctx->lpr;
    g.super_rounds = n_super;
    g.tail_only = tail_only;
    g.regular_k = reg_key;
    g.smem_prices = ctx->tail_smem_prices;
    g.tail_smem_bytes = ctx->tail_smem_bytes;
    g.tail_max = solve_tail_max(ctx, forward);
    g.khosla_phases = khosla_phases(ctx);
    g.l2_bytes = ctx->l2_bytes;
    g.vals16 = narrow_scan_ptr(ctx);
    g.learned_wide = lw;
    g.own_max = solve_own_max(ctx, forward);
    g.zero_first = zero_first;
    *out = g.exec;
    return SLA_OK;
}

// Dynamic shared memory a tail launch may use: the 227 KB of an sm_100 CTA minus the kernel's static arrays.
constexpr uint32_t kTailDynSmemMax = 232448u - 2048u;

// Decide what the tail engine mirrors in shared memory for the current instance and how many bidders it accepts.
// Preference: prices + owners (u32, else u16 when person ids fit) > prices only > nothing, as long as at least
// min(256, requested) bidders still fit beside the mirrors.
void plan_tail(sla_ctx* c) {
    const uint32_t want = (uint32_t)c->opt_tail_max;
    uint32_t cap = 32;
    while (cap < want) cap <<= 1;
    bool sp = false;
    uint32_t om = 0;
    if (c->has_csr && want > 0) {
        const uint32_t M = c->n_cols, N = c->n_rows;
        const uint32_t min_cap = cap < 256u ? cap : 256u;
        struct Cand { bool sp; uint32_t om; bool ok; };
        const Cand cands[4] = {
            {true, 1u, c->opt_smem_prices && c->opt_smem_owners},
            {true, 2u, c->opt_smem_prices && c->opt_smem_owners && N <= 65535u},
            {true, 0u, c->opt_smem_prices != 0},
            {false, 0u, true},
        };
        bool chosen = false;
        for (const Cand& k : cands) {
            if (!k.ok) continue;
            const uint64_t mirror = (uint64_t)M * ((k.sp ? 8u : 0u) + (k.om == 1u ? 4u : (k.om == 2u ? 2u : 0u)));
            if (mirror > kTailDynSmemMax) continue;
            for (uint32_t cc = cap; cc >= min_cap && !chosen; cc >>= 1) {
                if (tail_smem_layout(k.sp, k.om, M, cc).total <= kTailDynSmemMax) {
                    sp = k.sp; om = k.om; cap = cc; chosen = true;
                }
            }
            if (chosen) break;
        }
    }
    c->tail_smem_prices = sp;
    c->tail_own_mode = om;
    c->tail_cap = cap;
    c->tail_max_eff = want < cap ? want : cap;
    // Without the shared-memory price mirror (large M) a round of several hundred bidders is a few serial passes of
    // dependent global loads in the one CTA: the grid-wide pair is faster down to ~500 bidders (cfg3: 0.197 -> 0.187 ms).
    // An explicit "tail_max" option is taken as given.
    if (!sp && !c->tail_max_user && c->tail_max_eff > 512u) c->tail_max_eff = 512u;
    c->tail_smem_bytes = tail_smem_layout(sp, om, c->has_csr ? c->n_cols : 0u, cap).total;
}

double get_toleration_host(double c) {
    // reference src/solver.rs:144-146: 1 / 2^(53 - (log2(c + 1e-7) as u32)), saturating cast
    const double l = std::log2(c + 1e-7);
    uint32_t li = !(l > 0.0) ? 0u : (l >= 4294967295.0 ? 4294967295u : (uint32_t)l);
    const uint32_t e = 53u - li;
    const uint64_t pw = e < 64 ? ((uint64_t)1 << e) : 0;
    return 1.0 / (double)pw;
}

uint32_t person_bits(uint32_t n_rows) {
    uint32_t m = n_rows > 1 ? n_rows - 1 : 1;
    uint32_t b = 0;
    while (m) { ++b; m >>= 1; }
    return b;
}

int poll_state(sla_ctx* ctx) {
    CU(cudaMemcpyAsync(ctx->h_state, ctx->d_state, sizeof(DevState), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SLA_OK;
}

int finish_csr(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, uint64_t nnz, bool build_mirror = false, bool prestats = false);

// The solve driver shared by both algorithms.
int solve_common(sla_ctx* ctx, int algo, int maximize, double eps_in, double start_eps_in, uint32_t max_iterations,
                 uint32_t* h_p2o, uint32_t* h_o2p, double* h_prices, sla_stats* stats) {
    if (!ctx) return SLA_ERR_INVALID;
    NvtxRange nvtx_solve(algo == SLA_ALGO_FORWARD ? "sla_forward_solve" : "sla_khosla_solve");
    WorkerJoinGuard join_guard{ctx};
    if (!ctx->has_csr) return fail(ctx, SLA_ERR_STATE, "solve called before a CSR was uploaded");
    CU(cudaSetDevice(ctx->device));
    const bool forward = (algo == SLA_ALGO_FORWARD);
    const uint32_t N = ctx->n_rows, M = ctx->n_cols;

    // ---- sign normalisation, reference src/solver.rs:207-216 (the device keeps the uploaded values and a sign) ----
    const double first_eff = ctx->dev_sign > 0 ? ctx->first_value : -ctx->first_value;
    const bool positive = first_eff >= 0.0;
    const bool flip = (maximize != 0) != positive;
    if (flip) ctx->dev_sign = -ctx->dev_sign;
    const double w_min = ctx->dev_sign > 0 ? ctx->v_min : -ctx->v_max;
    const double w_max = ctx->dev_sign > 0 ? ctx->v_max : -ctx->v_min;

    DevState s;
    memset(&s, 0, sizeof s);
    s.qlen[0] = N;
    s.cur = 0;
    s.identity = 1;
    s.zero_prices = 1;
    s.algo = forward ? ALGO_FORWARD : ALGO_KHOSLA;
    s.pbits = person_bits(N);
    s.tail_max = solve_tail_max(ctx, forward);
    s.tail_own_max = solve_own_max(ctx, forward);
    s.own_mode = ctx->tail_own_mode;
    s.tail_cap = ctx->tail_cap;
    s.skip_zero = (uint32_t)ctx->opt_skip_zero;
    s.sign_flip = ctx->dev_sign < 0 ? 0x80000000u : 0u;
    s.n_rows = N;
    s.n_cols = M;
    s.safety_rounds_left = 1ull << 40;
    s.tail_round_cap = 1u << 19;
    s.regular_k = use_regular(ctx) ? ctx->regular_k : 0u;
    if (!forward) {
        // reference src/ksparse.rs:160-181
        const double m = (double)M;
        s.eps = std::isnan(eps_in) ? 1.0 / m : eps_in;
        s.threshold = (m / 2.0) * (w_max - w_min + s.eps);
        s.max_iterations = 0xFFFFFFFFu;
        s.target_eps = s.eps;
        // square instance: the rounds start at c/2 and come down to the caller's eps by factors of 0.15 (same factor
        // and start as the Forward solver, symmetric.rs:189, 270); see finish_if_possible
        const double c = std::fmax(std::fabs(w_min), std::fabs(w_max));
        if (khosla_phases(ctx) && c / 2.0 > s.eps) { s.kscale = 1u; s.eps = c / 2.0; }
    } else {
        // reference src/symmetric.rs:229-273
        const double target = std::isnan(eps_in) ? 1.0 / (double)N : eps_in;
        s.target_eps = target;
        s.max_iterations = max_iterations ? max_iterations : 100000u;
        const double c = std::fmax(std::fabs(w_min), std::fabs(w_max));
        s.tol = get_toleration_host(c);
        bool start_opt = !std::isnan(start_eps_in) ? (start_eps_in < target) : false;
        if (N != M) {
            start_opt = true;
            s.eps = target - 2.220446049250313e-16;
        } else {
            s.eps = !std::isnan(start_eps_in) ? start_eps_in : c / 2.0;
        }
        s.start_opt = start_opt ? 1u : 0u;
    }
    // every eps of the solve is >= 0 (later phases only multiply by 0.15): prices stay >= 0, the gathering scan may prune
    s.prune_ok = (ctx->opt_prune && s.eps >= 0.0 && s.target_eps >= 0.0) ? 1u : 0u;
    *ctx->h_state = s;

    const Params p = make_params(ctx);
    uint32_t launches = 0, graph_launches = 0;
    ctx->profile.clear();
    ctx->has_solution = false;

    // The initialisation of prices / owners / assignment (solver.rs:218-229) runs BEHIND the first bid scan when that
    // scan does not read them (all prices zero, wide first round); otherwise up front.
    const bool late_init = ctx->opt_skip_zero != 0 && !is_tail_only(ctx, forward);
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_state, ctx->h_state, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    if (late_init) {
        if (ctx->best_dirty || ctx->opt_prezero_best)
            CU(cudaMemsetAsync(ctx->d_best, 0, (size_t)M * sizeof(unsigned long long), ctx->stream));
    } else {
        init_solve_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, N, M, ctx->best_dirty ? 1 : 0);
        launches += 1;
    }
    ctx->best_dirty = true;   // until the solve ends at a round boundary

    const bool use_graph = ctx->opt_graph && !ctx->opt_profile;
    const size_t dl_o2p = h_p2o ? (((size_t)N * 4 + 15) & ~(size_t)15) : 0;
    const size_t dl_prices = dl_o2p + (h_o2p ? (((size_t)M * 4 + 15) & ~(size_t)15) : 0);
    const size_t dl_bytes = dl_prices + (h_prices ? (size_t)M * 8 : 0);
    constexpr size_t kSmallDownloadBytes = (size_t)1 << 20;
    // also with nothing to download (resident solves): the events are then recorded behind the graph launch and the
    // poll of the control block is the only synchronisation of the solve
    const bool small_dl = use_graph && ctx->opt_small_path && dl_bytes <= kSmallDownloadBytes;
    if (small_dl && dl_bytes > 0 && !ctx->h_dl) {
        CU(cudaMallocHost((void**)&ctx->h_dl, kSmallDownloadBytes));
        ctx->h_dl_cap = kSmallDownloadBytes;
    }
    bool done = false;
    const auto t_start = std::chrono::steady_clock::now();
    auto timed_out = [&]() {
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > ctx->opt_timeout_s;
    };
    if (use_graph) {
        cudaGraphExec_t exec_first = nullptr, exec_next = nullptr;
        int rc = get_graph(ctx, forward, ctx->opt_skip_zero != 0, true, &exec_first);
        if (rc) return rc;
        bool first = true;
        while (!done) {
            cudaGraphExec_t exec = exec_first;
            const bool first_graph = first;
            if (!first) {
                if (!exec_next && (rc = get_graph(ctx, forward, false, false, &exec_next))) return rc;
                exec = exec_next;
            }
            first = false;
            {
                NvtxRange nvtx_graph(first_graph ? "super-rounds graph (first)" : "super-rounds graph (continuation)");
                CU(cudaGraphLaunch(exec, ctx->stream));
            }
            graph_launches += 1;
            if (small_dl) {
                // small results ride behind every graph launch into page-locked staging: when the poll below reports
                // `done`, they are already on the host (one synchronisation per launch instead of three)
                CU(cudaEventRecord(ctx->ev[1], ctx->stream));
                if (h_p2o) CU(cudaMemcpyAsync(ctx->h_dl, ctx->d_p2o, (size_t)N * 4, cudaMemcpyDeviceToHost, ctx->stream));
                if (h_o2p) CU(cudaMemcpyAsync(ctx->h_dl + dl_o2p, ctx->d_o2p, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
                if (h_prices) CU(cudaMemcpyAsync(ctx->h_dl + dl_prices, ctx->d_prices, (size_t)M * 8, cudaMemcpyDeviceToHost, ctx->stream));
                CU(cudaEventRecord(ctx->ev[2], ctx->stream));
            }
            launches += (uint32_t)graph_kernel_count(ctx, forward, first_graph, late_init);
            if ((rc = poll_state(ctx))) return rc;
            done = ctx->h_state->done != 0;
            if (!done && timed_out()) return fail(ctx, SLA_ERR_STATE, "solve exceeded the wall-clock guard (timeout_s)");
        }
    } else {
        // Host-driven loop: one super-round per iteration, state polled after each.  With "profile" the wide
        // kernels and the tail engine are bracketed by CUDA events on the solve stream.
        DevState prev = s;
        bool first = true;
        while (!done) {
            const bool wide = !prev.done && prev.qlen[prev.cur] > prev.tail_max;
            // "profile_repeat" (development): the scan is idempotent (same slots, same maxima), so it can be launched
            // several times between the two events to separate its duration from the event overhead
            const int reps = (ctx->opt_profile && wide && first) ? ctx->opt_profile_repeat : 1;
            bool bracketed = false;
            if (ctx->opt_profile && ctx->opt_profile_graph && wide) {
                // The bracket as ONE graph launch -- fence kernel, event-record node, the scan, event-record node -- so
                // that the two time stamps are taken by the device's own node-to-node sequencing (what a scan inside the
                // solve graph sees) and not across two host-side launch latencies.  Any failure falls back to the eager bracket.
                cudaGraph_t pg = nullptr;
                if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    profile_fence_kernel<<<1, 32, 0, ctx->stream>>>();
                    cudaError_t e1 = cudaEventRecordWithFlags(ctx->ev[1], ctx->stream, cudaEventRecordExternal);
                    for (int r = 0; r < reps; ++r) launch_one(ctx, p, 0, first && ctx->opt_skip_zero != 0);
                    cudaError_t e2 = cudaEventRecordWithFlags(ctx->ev[2], ctx->stream, cudaEventRecordExternal);
                    cudaError_t e3 = cudaStreamEndCapture(ctx->stream, &pg);
                    if (e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess && pg) {
                        window_on_kernel_nodes(ctx, pg);
                        if (ctx->profile_exec) { cudaGraphExecDestroy(ctx->profile_exec); ctx->profile_exec = nullptr; }
                        if (cudaGraphInstantiate(&ctx->profile_exec, pg, 0) == cudaSuccess &&
                            cudaGraphLaunch(ctx->profile_exec, ctx->stream) == cudaSuccess)
                            bracketed = true;
                    }
                    if (pg) cudaGraphDestroy(pg);
                }
                if (!bracketed) {
                    cudaGetLastError();
                    if (!ctx->profile_graph_failed) fprintf(stderr, "[sla] profile: graph bracket unavailable, using the eager bracket\n");
                    ctx->profile_graph_failed = true;
                }
            }
            if (!bracketed) {
                if (ctx->opt_profile) {
                    profile_fence_kernel<<<1, 32, 0, ctx->stream>>>();
                    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
                }
                for (int r = 0; r < reps; ++r) launch_one(ctx, p, 0, first && ctx->opt_skip_zero != 0);
                if (ctx->opt_profile) CU(cudaEventRecord(ctx->ev[2], ctx->stream));
            }
            const bool late_now = first && late_init;
            if (late_now) {
                first_assign_objects_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p);
                launches += 1;
            }
            first = false;
            launch_one(ctx, p, 1, late_now);
            if (ctx->opt_profile) CU(cudaEventRecord(ctx->ev[3], ctx->stream));
            launches += 2;
            if (ctx->opt_profile) {
                int rc = poll_state(ctx);   // state after the wide pair and its control step A
                if (rc) return rc;
                if (wide) {
                    sla_round_profile r;
                    memset(&r, 0, sizeof r);
                    r.round = (uint32_t)prev.rounds;
                    r.engine = 0;
                    r.bidders = prev.qlen[prev.cur];
                    r.rounds_covered = 1;
                    r.arcs = prev.regular_k ? (uint64_t)r.bidders * prev.regular_k : ctx->h_state->bid_arcs - prev.bid_arcs;
                    cudaEventElapsedTime(&r.bid_ms, ctx->ev[1], ctx->ev[2]);
                    cudaEventElapsedTime(&r.assign_ms, ctx->ev[2], ctx->ev[3]);   // first round: object pass + person pass
                    ctx->profile.push_back(r);
                }
                CU(cudaEventRecord(ctx->ev[1], ctx->stream));
            }
            if (use_mid(ctx)) { launch_one(ctx, p, 5); launches += 1; }
            launch_one(ctx, p, 2);
            launches += 1;
            if (ctx->opt_profile) {
                CU(cudaEventRecord(ctx->ev[2], ctx->stream));
                const DevState mid = *ctx->h_state;
                int rc = poll_state(ctx);
                if (rc) return rc;
                if (ctx->h_state->tail_rounds > mid.tail_rounds) {
                    sla_round_profile r;
                    memset(&r, 0, sizeof r);
                    r.round = (uint32_t)mid.rounds;            // (control step A of the wide round has already run)
                    r.engine = 1;
                    r.bidders = mid.qlen[mid.cur];
                    r.rounds_covered = (uint32_t)(ctx->h_state->tail_rounds - mid.tail_rounds);
                    r.arcs = ctx->h_state->bid_arcs - mid.bid_arcs;
                    cudaEventElapsedTime(&r.bid_ms, ctx->ev[1], ctx->ev[2]);
                    ctx->profile.push_back(r);
                }
            }
            if (forward) {
                launch_one(ctx, p, 3);
                launch_one(ctx, p, 4);
                launches += 2;
            } else if (khosla_phases(ctx)) {
                launch_one(ctx, p, 4);
                launches += 1;
            }
            int rc = poll_state(ctx);
            if (rc) return rc;
            prev = *ctx->h_state;
            done = prev.done != 0;
            if (!done && timed_out()) return fail(ctx, SLA_ERR_STATE, "solve exceeded the wall-clock guard (timeout_s)");
        }
    }
    if (!small_dl) CU(cudaEventRecord(ctx->ev[1], ctx->stream));

    const DevState f = *ctx->h_state;
    if (f.safety_rounds_left <= 1) return fail(ctx, SLA_ERR_STATE, "safety round limit reached");
    ctx->best_dirty = false;
    ctx->has_solution = true;
    if (!forward && !khosla_phases(ctx)) ctx->learned_wide = (uint32_t)(f.wide_rounds < 65u ? f.wide_rounds : 0u);

    if (small_dl) {
        // already downloaded behind the last graph launch (events 1 and 2 recorded there, completed by the poll's sync)
        if (h_p2o) memcpy(h_p2o, ctx->h_dl, (size_t)N * 4);
        if (h_o2p) memcpy(h_o2p, ctx->h_dl + dl_o2p, (size_t)M * 4);
        if (h_prices) memcpy(h_prices, ctx->h_dl + dl_prices, (size_t)M * 8);
    } else {
        NvtxRange nvtx_dl("download solution");
        if (h_p2o) CU(cudaMemcpyAsync(h_p2o, ctx->d_p2o, (size_t)N * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (h_o2p) CU(cudaMemcpyAsync(h_o2p, ctx->d_o2p, (size_t)M * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (h_prices) CU(cudaMemcpyAsync(h_prices, ctx->d_prices, (size_t)M * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaEventRecord(ctx->ev[2], ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }

    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->nreductions = f.nreductions;
        stats->optimal_soln_found = f.optimal;
        stats->eps = f.eps;
        stats->rounds = f.rounds;
        stats->bids = f.bids;
        stats->bid_arcs = f.bid_arcs;
        stats->dropped = f.dropped;
        stats->values_negated = flip ? 1u : 0u;
        stats->wide_rounds = f.wide_rounds;
        stats->tail_rounds = f.tail_rounds;
        stats->cluster_rounds = f.cluster_rounds;
        stats->restarts = (s.kscale && !f.kscale) ? 1u : 0u;   // the eps-schedule was abandoned (finish_if_possible)
        stats->kernel_launches = launches;
        stats->graph_launches = graph_launches;
        if (forward) {
            stats->nits = f.nits;
            stats->num_unassigned = f.qlen[f.cur];
        } else {
            stats->nits = (uint32_t)f.bids;
            stats->num_unassigned = f.dropped;
        }
        cudaEventElapsedTime(&stats->ms_solve, ctx->ev[0], ctx->ev[1]);
        cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[2]);
    }
    return SLA_OK;
}

// After the three CSR arrays are resident: statistics + validation (solver.rs:232-243).
// Resets the statistics block on the device (stream-ordered; the page-locked source is rewritten with the same bytes only).
int csr_stats_reset(sla_ctx* ctx) {
    DevCsrStats init;
    init.min_key = ~0ull; init.max_key = 0ull; init.bad_cols = 0; init.bad_rows = 0; init.irregular_rows = 0; init.not_u16 = 0;
    *ctx->h_csr_stats = init;
    CU(cudaMemcpyAsync(ctx->d_csr_stats, ctx->h_csr_stats, sizeof(DevCsrStats), cudaMemcpyHostToDevice, ctx->stream));
    return SLA_OK;
}

// `prestats`: the statistics were accumulated on the way up (u16 uploads: csr_row_stats_kernel + widen_u16_stats_kernel
// behind the copies); only their read-back is left.
int finish_csr(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, uint64_t nnz, bool build_mirror, bool prestats) {
    if (!prestats) {
        int rc0 = csr_stats_reset(ctx);
        if (rc0) return rc0;
        csr_stats_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(ctx->d_row_ptr, ctx->d_cols, ctx->d_vals, num_rows,
                                                                          num_cols, nnz, ctx->d_csr_stats);
    }
    CU(cudaMemcpyAsync(ctx->h_csr_stats, ctx->d_csr_stats, sizeof(DevCsrStats), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_partial, ctx->d_vals, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_scratch, ctx->d_row_ptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->h_scratch + 1, ctx->d_row_ptr + num_rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->has_csr = false;
    ctx->vals16_valid = false;       // upload_large sets it again when the values it just sent went up as u16
    if (ctx->h_scratch[0] != 0 || (uint64_t)ctx->h_scratch[1] != nnz)
        return fail(ctx, SLA_ERR_INVALID, "row_ptr[0] must be 0 and row_ptr[num_rows] must equal nnz");
    if (ctx->h_csr_stats->bad_rows) return fail(ctx, SLA_ERR_INVALID, "row extents are not monotone");
    if (ctx->h_csr_stats->bad_cols) return fail(ctx, SLA_ERR_INVALID, "column index out of range (>= num_cols)");
    ctx->n_rows = num_rows;
    ctx->n_cols = num_cols;
    ctx->nnz = nnz;
    ctx->v_min = order_key_to_f64_host(ctx->h_csr_stats->min_key);
    ctx->v_max = order_key_to_f64_host(ctx->h_csr_stats->max_key);
    ctx->first_value = ctx->h_partial[0];
    ctx->dev_sign = 1;
    ctx->lpr = pick_lpr(nnz, num_rows);
    ctx->regular_k = 0;
    if (ctx->h_csr_stats->irregular_rows == 0 && nnz == (uint64_t)num_rows * (nnz / num_rows) && (nnz / num_rows) % 8 == 0) {
        ctx->regular_k = (uint32_t)(nnz / num_rows);
        int l = 1;
        while (l < 32 && (uint32_t)(l * 8) < ctx->regular_k) l *= 2;
        ctx->lpr8 = l;
    }
    // all values are u16 integers and no u16 copy came with the upload (CSR generated in HBM / handed over in device
    // memory): build the mirror the uniform-degree scans read (6 instead of 12 bytes per arc)
    if (build_mirror && ctx->regular_k && ctx->opt_narrow_scan && ctx->h_csr_stats->not_u16 == 0 && nnz >= ((uint64_t)1 << 20)) {
        const size_t need = (size_t)nnz * 2u;
        if (need > ctx->d_narrow_cap) {
            if (ctx->d_narrow) { cudaFree(ctx->d_narrow); ctx->d_narrow = nullptr; ctx->d_narrow_cap = 0; }
            if (cudaMalloc(&ctx->d_narrow, need) == cudaSuccess) ctx->d_narrow_cap = need; else cudaGetLastError();
        }
        if (need <= ctx->d_narrow_cap) {
            narrow_mirror_kernel<<<ctx->num_sms * 8, kWideThreads, 0, ctx->stream>>>(ctx->d_vals, (uint16_t*)ctx->d_narrow, (size_t)nnz);
            ctx->vals16_valid = true;
        }
    }
    ctx->has_csr = true;
    ctx->has_solution = false;
    ctx->learned_wide = 0;
    plan_tail(ctx);
    apply_l2_policy(ctx);
    return SLA_OK;
}

// ---- small instances: no device round trip on the way up ----------------------------------------------------------
// The statistics csr_stats_kernel would compute (value range, column bound, monotone extents, uniform degree) are taken
// on the host in the same pass that copies the three arrays into one page-locked block; the copies to the device are
// only enqueued.  With `negate_host` the caller's values are negated in place in that pass (solver.rs:214-216); the
// device receives the original values, as with sla_upload_csr_negating.  The caller's arrays are not referenced after
// the call returns.
constexpr size_t kSmallCsrBytes = (size_t)1 << 20;

bool small_csr(uint32_t num_rows, uint64_t nnz) { return ((size_t)num_rows + 1) * 4 + (size_t)nnz * 12 <= kSmallCsrBytes; }

int small_block(sla_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->h_small_cap) return SLA_OK;
    if (ctx->small_pending) { CU(cudaEventSynchronize(ctx->ev_small)); ctx->small_pending = false; }
    if (ctx->h_small) { cudaFreeHost(ctx->h_small); ctx->h_small = nullptr; ctx->h_small_cap = 0; }
    size_t cap = (size_t)64 << 10;
    while (cap < bytes) cap <<= 1;
    CU(cudaMallocHost((void**)&ctx->h_small, cap));
    ctx->h_small_cap = cap;
    if (!ctx->ev_small) CU(cudaEventCreateWithFlags(&ctx->ev_small, cudaEventDisableTiming));
    return SLA_OK;
}

// Copy `values` into the staging block, take their range, optionally negate the caller's copy in place.  Returns true
// when a zero or a NaN was seen: the caller then recomputes the range in the exact total order of csr_stats_kernel.
bool stage_values(double* values, double* stage, uint64_t nnz, bool negate, double* omn, double* omx) {
    uint64_t g = 0;
    double mn = values[0], mx = values[0];
    bool special = false;
#if defined(__SSE2__)
    if (nnz >= 8) {
        __m128d mn0 = _mm_set1_pd(mn), mn1 = mn0, mx0 = mn0, mx1 = mn0;
        __m128d sp = _mm_setzero_pd();
        const __m128d zero = _mm_setzero_pd(), sign = _mm_set1_pd(-0.0);
        for (; g + 4 <= nnz; g += 4) {
            const __m128d a = _mm_loadu_pd(values + g), b = _mm_loadu_pd(values + g + 2);
            _mm_storeu_pd(stage + g, a);
            _mm_storeu_pd(stage + g + 2, b);
            mn0 = _mm_min_pd(a, mn0); mx0 = _mm_max_pd(a, mx0);
            mn1 = _mm_min_pd(b, mn1); mx1 = _mm_max_pd(b, mx1);
            sp = _mm_or_pd(sp, _mm_or_pd(_mm_cmpeq_pd(a, zero), _mm_cmpunord_pd(a, a)));
            sp = _mm_or_pd(sp, _mm_or_pd(_mm_cmpeq_pd(b, zero), _mm_cmpunord_pd(b, b)));
            if (negate) {
                _mm_storeu_pd(values + g, _mm_xor_pd(a, sign));
                _mm_storeu_pd(values + g + 2, _mm_xor_pd(b, sign));
            }
        }
        mn0 = _mm_min_pd(mn0, mn1);
        mx0 = _mm_max_pd(mx0, mx1);
        double t[2];
        _mm_storeu_pd(t, mn0); mn = t[0] < t[1] ? t[0] : t[1];
        _mm_storeu_pd(t, mx0); mx = t[0] > t[1] ? t[0] : t[1];
        special = _mm_movemask_pd(sp) != 0;
    }
#endif
    for (; g < nnz; ++g) {
        const double v = values[g];
        stage[g] = v;
        mn = v < mn ? v : mn;
        mx = v > mx ? v : mx;
        special |= !(v == v) || v == 0.0;
        if (negate) values[g] = -v;
    }
    *omn = mn;
    *omx = mx;
    return special;
}

// Copy the column indices into the staging block; true when one of them is >= num_cols.
bool stage_cols(const uint32_t* cols, uint32_t* stage, uint64_t nnz, uint32_t num_cols) {
    uint64_t g = 0;
    bool bad = false;
#if defined(__SSE2__)
    const __m128i bias = _mm_set1_epi32((int)0x80000000u), lim = _mm_set1_epi32((int)((num_cols - 1u) ^ 0x80000000u));
    __m128i acc = _mm_setzero_si128();
    for (; g + 8 <= nnz; g += 8) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(cols + g)), b = _mm_loadu_si128((const __m128i*)(cols + g + 4));
        _mm_storeu_si128((__m128i*)(stage + g), a);
        _mm_storeu_si128((__m128i*)(stage + g + 4), b);
        acc = _mm_or_si128(acc, _mm_or_si128(_mm_cmpgt_epi32(_mm_xor_si128(a, bias), lim), _mm_cmpgt_epi32(_mm_xor_si128(b, bias), lim)));
    }
    bad = _mm_movemask_epi8(acc) != 0;
#endif
    for (; g < nnz; ++g) {
        stage[g] = cols[g];
        bad |= cols[g] >= num_cols;
    }
    return bad;
}

int upload_small(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t* row_ptr, const uint32_t* column_indices,
                 double* values, uint64_t nnz, bool negate_host) {
    NvtxRange nvtx_up("upload CSR (small: host statistics + staged H2D)");
    const size_t b_rp = (((size_t)num_rows + 1) * 4 + 15) & ~(size_t)15, b_cols = ((size_t)nnz * 4 + 15) & ~(size_t)15;
    int rc = small_block(ctx, b_rp + b_cols + (size_t)nnz * 8);
    if (rc) return rc;
    if (ctx->small_pending) { CU(cudaEventSynchronize(ctx->ev_small)); ctx->small_pending = false; }
    uint32_t* s_rp = reinterpret_cast<uint32_t*>(ctx->h_small);
    uint32_t* s_cols = reinterpret_cast<uint32_t*>(ctx->h_small + b_rp);
    double* s_vals = reinterpret_cast<double*>(ctx->h_small + b_rp + b_cols);
    ctx->has_csr = false;
    ctx->vals16_valid = false;
    // extents
    if (row_ptr[0] != 0 || (uint64_t)row_ptr[num_rows] != nnz)
        return fail(ctx, SLA_ERR_INVALID, "row_ptr[0] must be 0 and row_ptr[num_rows] must equal nnz");
    const uint32_t k0 = row_ptr[1] - row_ptr[0];
    uint32_t bad_r = 0, irr = 0;
    s_rp[0] = row_ptr[0];
    for (uint32_t i = 0; i < num_rows; ++i) {
        const uint32_t a = row_ptr[i], b = row_ptr[i + 1];
        s_rp[i + 1] = b;
        bad_r |= (b < a || (uint64_t)b > nnz) ? 1u : 0u;
        irr |= (b - a != k0) ? 1u : 0u;
    }
    if (bad_r) return fail(ctx, SLA_ERR_INVALID, "row extents are not monotone");
    // Columns, values: one vectorised pass each (copy + bound check, copy + range [+ in-place negation]).  Every piece
    // goes onto the wire as soon as it is staged, so that the DMA engine works while the host stages the next one: the
    // upload ends one piece's transfer after the last pass instead of the whole block's (the values travel in two
    // pieces from 128 KB up).  Once a copy is in flight the staging block is guarded by ev_small, also on the error paths.
    auto in_flight = [&]() {
        cudaEventRecord(ctx->ev_small, ctx->stream);
        ctx->small_pending = true;
    };
    CU(cudaMemcpyAsync(ctx->d_row_ptr, s_rp, ((size_t)num_rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (stage_cols(column_indices, s_cols, nnz, num_cols)) {
        in_flight();
        return fail(ctx, SLA_ERR_INVALID, "column index out of range (>= num_cols)");
    }
    {
        const cudaError_t e = cudaMemcpyAsync(ctx->d_cols, s_cols, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { in_flight(); CU(e); }
    }
    const double first = values[0];
    double vmin = 0.0, vmax = 0.0;
    bool special = false;
    {
        constexpr uint64_t kPieceMin = (uint64_t)16 << 10;      // values (128 KB): below twice this, one piece
        const uint64_t half = nnz >= 2 * kPieceMin ? ((nnz / 2 + 3) & ~(uint64_t)3) : nnz;
        for (uint64_t lo = 0; lo < nnz; lo += half) {
            const uint64_t cnt = (lo + half < nnz) ? half : nnz - lo;
            double mn, mx;
            special |= stage_values(values + lo, s_vals + lo, cnt, negate_host, &mn, &mx);
            if (lo == 0) { vmin = mn; vmax = mx; }
            else { vmin = mn < vmin ? mn : vmin; vmax = mx > vmax ? mx : vmax; }
            const cudaError_t e = cudaMemcpyAsync(ctx->d_vals + lo, s_vals + lo, (size_t)cnt * 8, cudaMemcpyHostToDevice, ctx->stream);
            if (e != cudaSuccess) { in_flight(); CU(e); }
        }
    }
    in_flight();
    if (special) {
        // zeros or NaNs present: the range in csr_stats_kernel's total order (-0.0 < +0.0, NaNs at the ends), from the
        // staged copy of the original values
        unsigned long long kmin = ~0ull, kmax = 0ull;
        for (uint64_t g = 0; g < nnz; ++g) {
            unsigned long long u;
            __builtin_memcpy(&u, s_vals + g, 8);
            const unsigned long long k = u ^ ((u >> 63) ? ~0ull : (1ull << 63));
            kmin = k < kmin ? k : kmin;
            kmax = k > kmax ? k : kmax;
        }
        vmin = order_key_to_f64_host(kmin);
        vmax = order_key_to_f64_host(kmax);
    }
    ctx->n_rows = num_rows;
    ctx->n_cols = num_cols;
    ctx->nnz = nnz;
    ctx->v_min = vmin;
    ctx->v_max = vmax;
    ctx->first_value = first;
    ctx->dev_sign = 1;
    ctx->lpr = pick_lpr(nnz, num_rows);
    ctx->regular_k = 0;
    if (!irr && nnz == (uint64_t)num_rows * (nnz / num_rows) && (nnz / num_rows) % 8 == 0) {
        ctx->regular_k = (uint32_t)(nnz / num_rows);
        int l = 1;
        while (l < 32 && (uint32_t)(l * 8) < ctx->regular_k) l *= 2;
        ctx->lpr8 = l;
    }
    ctx->has_csr = true;
    ctx->has_solution = false;
    ctx->learned_wide = 0;
    plan_tail(ctx);
    apply_l2_policy(ctx);
    return SLA_OK;
}

int check_shape(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, uint64_t nnz) {
    // reference src/solver.rs:192-193 (init) and 232-240 (validate_input), for I = u32
    if (!(num_rows <= num_cols)) return fail(ctx, SLA_ERR_INVALID, "num_rows must be <= num_cols");
    if (!(num_rows < 0xFFFFFFFFu)) return fail(ctx, SLA_ERR_INVALID, "num_rows must be < u32::MAX");
    if (!(nnz > 0)) return fail(ctx, SLA_ERR_INVALID, "no arcs");
    if (!(num_rows > 0 && num_cols > 0)) return fail(ctx, SLA_ERR_INVALID, "num_rows and num_cols must be positive");
    if (!(nnz < 0xFFFFFFFFull)) return fail(ctx, SLA_ERR_INVALID, "number of arcs must be < u32::MAX");
    return SLA_OK;
}

// Large uploads.  `values` first tries to cross PCIe narrow: when every value survives a round trip through u16
// (non-negative integers below 65,536: the integer costs of the BASELINE configs) or through f32, the context's workers
// stage the narrow copy piece by piece while the column indices are on the wire, the pieces follow them, and
// widen_values_kernel restores the f64 array in HBM bit for bit (cfg3: 196 MB -> 100 MB over PCIe).  The first value
// that does not survive ends the attempt; the f64 values then go up as they are.  With `negate` the caller's copy is
// negated in place (solver.rs:214-216) by the same workers -- each piece right after it has been staged narrow (the
// original values are then no longer needed on the host) or, chunk by chunk, behind the f64 copies; the pool is drained before the next solve /
// upload / destroy call on this context returns.  The device always receives the ORIGINAL values.
int upload_large(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t* row_ptr, const uint32_t* column_indices,
                 double* values, uint64_t nnz, bool negate, int threads) {
    NvtxRange nvtx_up("upload CSR (large: narrow staging + H2D)");
    const size_t total = (size_t)nnz;
    const bool dbg = getenv("SLA_DEBUG_UPLOAD") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
    double t_first = 0, t_half = 0;
    ctx->last_upload_bytes = ((size_t)num_rows + 1) * 4 + total * 4;
    ctx->last_upload_value_bytes = 8;
    ctx->vals16_valid = false;
    const bool want_pool = negate || (ctx->opt_narrow_upload && total >= ((size_t)1 << 20));
    const bool pool_ok = want_pool && ctx->neg.start(ctx->device);
    if (negate && !pool_ok) return fail(ctx, SLA_ERR_CUDA, "could not start the host negation workers");
    NegPool& np = ctx->neg;
    if (threads < 1) {
        // default: the cores of the box divided by the ranks torchrun co-located on it
        unsigned hw = std::thread::hardware_concurrency();
        int lw = 1;
        if (const char* e = getenv("LOCAL_WORLD_SIZE")) lw = atoi(e) > 0 ? atoi(e) : 1;
        threads = (int)(hw ? hw : 4) / lw;
        if (const char* e = getenv("SLA_HOST_THREADS")) threads = atoi(e);
        if (threads < 1) threads = 1;
    }
    if (pool_ok && threads > (int)np.threads.size()) threads = (int)np.threads.size();

    // ---- tier: what do the first values look like? ---------------------------------------------------------------
    int tier = 0;
    if (ctx->opt_narrow_upload && pool_ok && total >= ((size_t)1 << 20)) {
        const size_t probe = 4096;
        std::vector<uint16_t> t16(probe);
        std::vector<float> t32(probe);
        if (NegPool::narrow_u16(values, t16.data(), 0, probe, false)) tier = 2;
        else if (NegPool::narrow_f32(values, t32.data(), 0, probe, false)) tier = 4;
    }
    if (tier) {
        const size_t need = total * (size_t)tier;
        if (need > ctx->h_narrow_cap) {
            if (ctx->h_narrow) { cudaFreeHost(ctx->h_narrow); ctx->h_narrow = nullptr; ctx->h_narrow_cap = 0; }
            if (cudaMallocHost(&ctx->h_narrow, need) == cudaSuccess) ctx->h_narrow_cap = need; else { cudaGetLastError(); tier = 0; }
        }
        if (tier && need > ctx->d_narrow_cap) {
            if (ctx->d_narrow) { cudaFree(ctx->d_narrow); ctx->d_narrow = nullptr; ctx->d_narrow_cap = 0; }
            if (cudaMalloc(&ctx->d_narrow, need) == cudaSuccess) ctx->d_narrow_cap = need; else { cudaGetLastError(); tier = 0; }
        }
    }

    // u16 tier: statistics and widening ride behind the copies (csr_row_stats_kernel behind the extents, one
    // widen_u16_stats_kernel behind every run of staged pieces), so that the upload ends a few microseconds after the
    // last piece has landed instead of a widening pass + a statistics pass over the whole CSR later (cfg3: ~0.11 ms)
    // The kernels run on a second stream, each behind an event of the copy it needs: the link is the second-longest
    // pole of the upload (100 MB at ~54 GB/s against 1.9 ms of host pass), kernels in between the copies on one stream
    // cost it 0.1 ms of bubbles.
    bool ride = tier == 2 && ctx->opt_upload_ride;
    if (ride && !ctx->stream_ride) {
        if (cudaStreamCreateWithFlags(&ctx->stream_ride, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&ctx->ev_ride, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            if (ctx->stream_ride) { cudaStreamDestroy(ctx->stream_ride); ctx->stream_ride = nullptr; }
            ride = false;
        }
    }
    auto ride_after_copy = [&]() -> cudaError_t {      // stream_ride continues behind what `stream` holds so far
        cudaError_t x = cudaEventRecord(ctx->ev_ride, ctx->stream);
        if (x == cudaSuccess) x = cudaStreamWaitEvent(ctx->stream_ride, ctx->ev_ride, 0);
        return x;
    };
    if (ride) { int rc0 = csr_stats_reset(ctx); if (rc0) return rc0; }
    CU(cudaMemcpyAsync(ctx->d_row_ptr, row_ptr, ((size_t)num_rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    if (ride) {
        CU(ride_after_copy());
        csr_row_stats_kernel<<<ctx->num_sms * 4, kWideThreads, 0, ctx->stream_ride>>>(ctx->d_row_ptr, num_rows, nnz, ctx->d_csr_stats);
    }
    CU(cudaMemcpyAsync(ctx->d_cols, column_indices, total * 4, cudaMemcpyHostToDevice, ctx->stream));

    bool narrowed = false;
    if (tier) {
        size_t piece = (size_t)1 << 16;                                    // 64 Ki values: 128 / 256 KB per copy
        while ((total + piece - 1) / piece > (size_t)NegPool::kMaxPieces) piece <<= 1;
        const int npieces = (int)((total + piece - 1) / piece);
        {
            std::lock_guard<std::mutex> lk(np.m);
            np.mode = NegPool::NARROW;
            np.values = values; np.total = total; np.stage = ctx->h_narrow; np.tier = tier; np.piece = piece; np.npieces = npieces;
            np.next_piece.store(0); np.narrow_fail.store(0);
            for (int q = 0; q < npieces; ++q) np.done[q].store(0, std::memory_order_relaxed);
            np.narrow_workers = threads;
            np.narrow_negate = negate;
            np.pending = threads;
            np.job_id += 1;
        }
        np.cv_job.notify_all();
        // pieces follow the column indices over PCIe in array order, each as soon as it is staged
        int sent = 0;
        constexpr int kMinRun = 16;
        cudaError_t ce = cudaSuccess;
        while (sent < npieces && ce == cudaSuccess) {
            if (np.done[sent].load(std::memory_order_acquire)) {
                int run = 1;                                              // one copy for every run of staged neighbours
                while (sent + run < npieces && np.done[sent + run].load(std::memory_order_acquire)) run += 1;
                if (run < kMinRun && sent + run < npieces) {                 // copies of >= 2 MB
                    if (np.narrow_fail.load(std::memory_order_acquire)) break;  // abandoned: no more pieces will follow
                    std::this_thread::yield();
                    continue;
                }
                const size_t lo = (size_t)sent * piece, end = (size_t)(sent + run) * piece, hi = end < total ? end : total;
                ce = cudaMemcpyAsync((char*)ctx->d_narrow + lo * tier, (const char*)ctx->h_narrow + lo * tier, (hi - lo) * tier,
                                     cudaMemcpyHostToDevice, ctx->stream);
                if (ride && ce == cudaSuccess && (ce = ride_after_copy()) == cudaSuccess) {
                    widen_u16_stats_kernel<<<ctx->num_sms * 8, kWideThreads, 0, ctx->stream_ride>>>(
                        (const uint16_t*)ctx->d_narrow, ctx->d_vals, ctx->d_cols, lo, hi, num_cols, ctx->d_csr_stats);
                    ce = cudaGetLastError();
                }
                if (dbg && sent == 0) t_first = ms_since();
                if (dbg && sent < npieces / 2 && sent + run >= npieces / 2) t_half = ms_since();
                sent += run;
            } else if (np.narrow_fail.load(std::memory_order_acquire)) {
                break;
            } else {
                std::this_thread::yield();
            }
        }
        if (dbg) fprintf(stderr, "[sla] narrow tier %d: first piece sent %.3f ms, half %.3f ms, all %.3f ms\n", tier, t_first, t_half, ms_since());
        np.wait();                                                        // the staging pass is over (or was abandoned)
        if (ride) {
            // `stream` continues behind the riding kernels (statistics complete, d_vals written; on abandonment the f64
            // copies below must not be overtaken by a kernel still widening an earlier run)
            cudaError_t x = cudaEventRecord(ctx->ev_ride, ctx->stream_ride);
            if (x == cudaSuccess) x = cudaStreamWaitEvent(ctx->stream, ctx->ev_ride, 0);
            if (ce == cudaSuccess) ce = x;
        }
        if (ce != cudaSuccess) return fail(ctx, SLA_ERR_CUDA, std::string("cudaMemcpyAsync: ") + cudaGetErrorString(ce));
        narrowed = sent == npieces && !np.narrow_fail.load();
        if (!narrowed && negate) {
            // abandoned late: the pieces staged so far were already negated in place -- restore them, the f64 path below
            // needs the original values on the host
            for (int q = 0; q < npieces; ++q)
                if (np.done[q].load(std::memory_order_acquire)) {
                    const size_t lo = (size_t)q * piece, hi = lo + piece < total ? lo + piece : total;
                    for (size_t i = lo; i < hi; ++i) values[i] = -values[i];
                }
        }
    }

    if (narrowed) {
        const int grid = ctx->num_sms * 8;
        if (ride) { /* widened run by run behind the copies */ }
        else if (tier == 2) widen_values_kernel<uint16_t><<<grid, kWideThreads, 0, ctx->stream>>>((const uint16_t*)ctx->d_narrow, ctx->d_vals, total);
        else           widen_values_kernel<float><<<grid, kWideThreads, 0, ctx->stream>>>((const float*)ctx->d_narrow, ctx->d_vals, total);
        ctx->last_upload_bytes += total * (size_t)tier;
        ctx->last_upload_value_bytes = (uint32_t)tier;
        // (with `negate` every piece was negated in place right after it was staged)
    } else if (negate) {
        // f64 values, chunk by chunk, each chunk followed by its event; worker w negates chunk w behind its copy
        size_t per = (total + (size_t)threads - 1) / (size_t)threads;
        per = (per + 511) & ~(size_t)511;                       // whole 4 KB pieces per worker
        const int nchunks = (int)((total + per - 1) / per);
        for (int w = 0; w < nchunks; ++w) {
            const size_t lo = (size_t)w * per, hi = lo + per < total ? lo + per : total;
            CU(cudaMemcpyAsync(ctx->d_vals + lo, values + lo, (hi - lo) * 8, cudaMemcpyHostToDevice, ctx->stream));
            CU(cudaEventRecord(np.ev[w], ctx->stream));
        }
        {
            std::lock_guard<std::mutex> lk(np.m);
            np.mode = NegPool::NEGATE_AFTER_COPY;
            np.values = values; np.chunk = per; np.total = total; np.nchunks = nchunks; np.pending = nchunks;
            np.job_id += 1;
        }
        np.cv_job.notify_all();
        ctx->last_upload_bytes += total * 8;
    } else {
        CU(cudaMemcpyAsync(ctx->d_vals, values, total * 8, cudaMemcpyHostToDevice, ctx->stream));
        ctx->last_upload_bytes += total * 8;
    }
    if (dbg) fprintf(stderr, "[sla] upload enqueued at %.3f ms\n", ms_since());
    int rc = finish_csr(ctx, num_rows, num_cols, nnz, false, narrowed && ride);
    if (dbg) fprintf(stderr, "[sla] upload complete (statistics read back) at %.3f ms\n", ms_since());
    if (rc) {
        // rejected (validate_input / column bound): no solve will follow, so the caller's values go back to what they were
        join_workers(ctx);
        if (negate) for (size_t i = 0; i < total; ++i) values[i] = -values[i];
    } else {
        ctx->vals16_valid = narrowed && tier == 2;   // d_narrow keeps the u16 image of d_vals until the next upload
    }
    return rc;
}

}  // namespace

// =============================================================================================================
// extern "C" boundary
// =============================================================================================================
extern "C" {

const char* sla_version(void) { return "sla_b200 0.1.0 (sm_100a)"; }

int sla_ctx_create(int device, size_t row_capacity, size_t col_capacity, size_t arc_capacity, sla_ctx** out) {
    if (!out) return SLA_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return SLA_ERR_NO_DEVICE;
    }
    if (device < 0) device = 0;
    if (device >= count) { g_create_error = "device index out of range"; return SLA_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return SLA_ERR_NO_DEVICE;
    }
    if (prop.major != 10) {
        g_create_error = "this library contains sm_100a code only; device is sm_" + std::to_string(prop.major * 10 + prop.minor);
        return SLA_ERR_NO_DEVICE;
    }
    sla_ctx* ctx = new sla_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (const char* t = getenv("SLA_TIMEOUT_S")) ctx->opt_timeout_s = atof(t);
    auto bail = [&](const char* what, cudaError_t err) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(err);
        sla_ctx_destroy(ctx);
        return SLA_ERR_CUDA;
    };
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (auto& ev : ctx->ev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaMallocHost((void**)&ctx->h_state, sizeof(DevState))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_csr_stats, sizeof(DevCsrStats))) != cudaSuccess) return bail("cudaMallocHost", e);
    if ((e = cudaMallocHost((void**)&ctx->h_scratch, 16 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMallocHost", e);
    {
        const int dyn = (int)kTailDynSmemMax;
        cudaError_t ea[12] = {
            cudaFuncSetAttribute(tail_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
            cudaFuncSetAttribute(tail_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn),
        };
        for (cudaError_t x : ea)
            if (x != cudaSuccess) return bail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", x);
        const int sdyn = (int)kStreamSmemBytes;
        cudaError_t eb[6] = {
            cudaFuncSetAttribute(bid_stream_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
            cudaFuncSetAttribute(bid_stream_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
            cudaFuncSetAttribute(bid_stream_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
            cudaFuncSetAttribute(bid_stream_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
            cudaFuncSetAttribute(bid_stream_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
            cudaFuncSetAttribute(bid_stream_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, sdyn),
        };
        for (cudaError_t x : eb)
            if (x != cudaSuccess) return bail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", x);
    }
    // blocks per SM of the widest kernel decide the persistent grid
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bid_wide_kernel<4>, kWideThreads, 0);
    if (occ < 1) occ = 1;
    if (occ > 8) occ = 8;
    ctx->grid_wide = ctx->num_sms * occ;
    if ((e = cudaMallocHost((void**)&ctx->h_partial, (size_t)ctx->grid_wide * sizeof(double))) != cudaSuccess)
        return bail("cudaMallocHost", e);
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->d_state, 1)) || (rc = dev_alloc(ctx, &ctx->d_csr_stats, 1)) ||
        (rc = dev_alloc(ctx, &ctx->d_scratch, 16)) || (rc = dev_alloc(ctx, &ctx->d_partial, (size_t)ctx->grid_wide)) ||
        (rc = ensure_capacity(ctx, row_capacity ? row_capacity : 1, col_capacity ? col_capacity : 1,
                              arc_capacity ? arc_capacity : 1))) {
        g_create_error = ctx->err;
        sla_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return SLA_OK;
}

void sla_batch_free(sla_ctx* ctx);
void sla_part_free(sla_ctx* ctx);
void sla_mesh_free(sla_ctx* ctx);

void sla_ctx_destroy(sla_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->neg.shutdown();
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    release_l2_policy(ctx);
    sla_batch_free(ctx);
    sla_part_free(ctx);
    sla_mesh_free(ctx);
    drop_graphs(ctx);
    if (ctx->profile_exec) cudaGraphExecDestroy(ctx->profile_exec);
    cudaFree(ctx->d_row_ptr); cudaFree(ctx->d_cols); cudaFree(ctx->d_vals); cudaFree(ctx->d_prices);
    cudaFree(ctx->d_p2o); cudaFree(ctx->d_o2p); cudaFree(ctx->d_best); cudaFree(ctx->d_queue[0]);
    cudaFree(ctx->d_queue[1]); cudaFree(ctx->d_slot_obj); cudaFree(ctx->d_slot_bid); cudaFree(ctx->d_state);
    cudaFree(ctx->d_csr_stats); cudaFree(ctx->d_scratch); cudaFree(ctx->d_partial);
    if (ctx->h_small) cudaFreeHost(ctx->h_small);
    if (ctx->h_dl) cudaFreeHost(ctx->h_dl);
    if (ctx->h_narrow) cudaFreeHost(ctx->h_narrow);
    if (ctx->d_narrow) cudaFree(ctx->d_narrow);
    if (ctx->ev_small) cudaEventDestroy(ctx->ev_small);
    if (ctx->ev_ride) cudaEventDestroy(ctx->ev_ride);
    if (ctx->stream_ride) cudaStreamDestroy(ctx->stream_ride);
    if (ctx->h_state) cudaFreeHost(ctx->h_state);
    if (ctx->h_csr_stats) cudaFreeHost(ctx->h_csr_stats);
    if (ctx->h_scratch) cudaFreeHost(ctx->h_scratch);
    if (ctx->h_partial) cudaFreeHost(ctx->h_partial);
    for (auto& ev : ctx->ev) if (ev) cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sla_last_error(const sla_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
void* sla_ctx_stream(sla_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
int sla_ctx_device(const sla_ctx* ctx) { return ctx ? ctx->device : -1; }

int sla_set_option(sla_ctx* ctx, const char* key, int64_t value) {
    if (!ctx || !key) return SLA_ERR_INVALID;
    const std::string k(key);
    if (k == "tail_max") {
        if (value < 0 || value > kTailCap) return fail(ctx, SLA_ERR_INVALID, "tail_max must be in [0, 1024]");
        ctx->opt_tail_max = (int)value;
        ctx->tail_max_user = true;
        plan_tail(ctx);
    } else if (k == "graph") {
        ctx->opt_graph = value ? 1 : 0;
    } else if (k == "zero_price_skip") {
        ctx->opt_skip_zero = value ? 1 : 0;
    } else if (k == "profile") {
        ctx->opt_profile = value ? 1 : 0;
    } else if (k == "smem_prices") {
        ctx->opt_smem_prices = value ? 1 : 0;
        plan_tail(ctx);
    } else if (k == "smem_owners") {
        ctx->opt_smem_owners = value ? 1 : 0;
        plan_tail(ctx);
    } else if (k == "khosla_scaling") {
        ctx->opt_khosla_scaling = value ? 1 : 0;
    } else if (k == "upload_ride") {
        ctx->opt_upload_ride = value ? 1 : 0;
    } else if (k == "cluster_engine") {
        ctx->opt_cluster_engine = value ? 1 : 0;
    } else if (k == "cluster_handover") {
        if (value < 1 || value > 512) return fail(ctx, SLA_ERR_INVALID, "cluster_handover must be in [1, 512]");
        ctx->opt_cluster_handover = (int)value;
    } else if (k == "prune_gather") {
        ctx->opt_prune = value ? 1 : 0;
    } else if (k == "learn_shape") {
        ctx->opt_learn_shape = value ? 1 : 0;
        ctx->learned_wide = 0;
    } else if (k == "mesh_tail") {
        ctx->opt_mesh_tail = value ? 1 : 0;
    } else if (k == "prezero_best") {
        ctx->opt_prezero_best = value ? 1 : 0;
    } else if (k == "regular") {
        ctx->opt_regular = value ? 1 : 0;
    } else if (k == "timeout_s") {
        ctx->opt_timeout_s = (double)value;
    } else if (k == "profile_graph") {
        ctx->opt_profile_graph = value ? 1 : 0;
    } else if (k == "profile_repeat") {
        if (value < 1 || value > 64) return fail(ctx, SLA_ERR_INVALID, "profile_repeat must be in [1, 64]");
        ctx->opt_profile_repeat = (int)value;
    } else if (k == "l2_persist") {
        ctx->opt_l2_persist = value ? 1 : 0;
        drop_graphs(ctx);
        apply_l2_policy(ctx);
    } else if (k == "narrow_upload") {
        ctx->opt_narrow_upload = value ? 1 : 0;
    } else if (k == "narrow_scan") {
        ctx->opt_narrow_scan = value ? 1 : 0;
    } else if (k == "small_path") {
        ctx->opt_small_path = value ? 1 : 0;
    } else if (k == "wide_first") {
        ctx->opt_wide_first = value ? 1 : 0;
    } else if (k == "stream_scan") {
        ctx->opt_stream_scan = value ? 1 : 0;
        drop_graphs(ctx);
    } else if (k == "super_rounds") {
        if (value < 1 || value > 64) return fail(ctx, SLA_ERR_INVALID, "super_rounds must be in [1, 64]");
        ctx->opt_super_rounds = (int)value;
        ctx->super_rounds_user = true;
    } else {
        return fail(ctx, SLA_ERR_INVALID, "unknown option: " + k);
    }
    return SLA_OK;
}

int sla_upload_csr(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t* row_ptr,
                   const uint32_t* column_indices, const double* values, uint64_t nnz) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!row_ptr || !column_indices || !values) return fail(ctx, SLA_ERR_INVALID, "null input array");
    join_workers(ctx);
    int rc = check_shape(ctx, num_rows, num_cols, nnz);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure_capacity(ctx, num_rows, num_cols, nnz))) return rc;
    if (ctx->opt_small_path && small_csr(num_rows, nnz)) {
        ctx->last_upload_bytes = ((size_t)num_rows + 1) * 4 + (size_t)nnz * 12;
        ctx->last_upload_value_bytes = 8;
        return upload_small(ctx, num_rows, num_cols, row_ptr, column_indices, const_cast<double*>(values), nnz, false);
    }
    return upload_large(ctx, num_rows, num_cols, row_ptr, column_indices, const_cast<double*>(values), nnz, false, 0);
}

// Same as sla_upload_csr, plus the in-place sign normalisation of AuctionSolver::init_solve (reference
// src/solver.rs:214-216) on the HOST values, done by the context's workers while the upload and the solve run (small
// instances: in the staging pass).  The pool is drained before the next solve / upload / destroy call on this context
// returns, so the caller observes the negated values when solve() returns -- exactly the reference's post-condition.
// The device holds the ORIGINAL values (the following solve reports values_negated == 1 as usual; the host must then
// not negate again).
int sla_upload_csr_negating(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t* row_ptr,
                            const uint32_t* column_indices, double* values, uint64_t nnz, int threads) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!row_ptr || !column_indices || !values) return fail(ctx, SLA_ERR_INVALID, "null input array");
    join_workers(ctx);
    int rc = check_shape(ctx, num_rows, num_cols, nnz);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure_capacity(ctx, num_rows, num_cols, nnz))) return rc;
    if (ctx->opt_small_path && small_csr(num_rows, nnz)) {
        ctx->last_upload_bytes = ((size_t)num_rows + 1) * 4 + (size_t)nnz * 12;
        ctx->last_upload_value_bytes = 8;
        return upload_small(ctx, num_rows, num_cols, row_ptr, column_indices, values, nnz, true);
    }
    return upload_large(ctx, num_rows, num_cols, row_ptr, column_indices, values, nnz, true, threads);
}

// Host helper behind the narrow upload, exposed for the wrappers' tests: tries to narrow values[0..n) to `tier` bytes
// each (2: u16, 4: f32) into `out`; returns 1 when every value survives the round trip bit for bit (then `out` holds the
// narrow copy and, with `negate`, `values` has been negated in place), 0 otherwise (`values` unchanged), -1 on bad
// arguments.  No device is involved.
int sla_host_narrow(double* values, size_t n, int tier, void* out, int negate) {
    if (!values || !out || (tier != 2 && tier != 4)) return -1;
    if (n == 0) return 1;
    const bool ok = tier == 2 ? NegPool::narrow_u16(values, (uint16_t*)out, 0, n, negate != 0)
                              : NegPool::narrow_f32(values, (float*)out, 0, n, negate != 0);
    return ok ? 1 : 0;
}

// Bytes the last upload moved from host to device and the width (2 / 4 / 8) the values crossed PCIe with.
int sla_last_upload(const sla_ctx* ctx, uint64_t* bytes, uint32_t* value_bytes) {
    if (!ctx) return SLA_ERR_INVALID;
    if (bytes) *bytes = ctx->last_upload_bytes;
    if (value_bytes) *value_bytes = ctx->last_upload_value_bytes;
    return SLA_OK;
}

// Width (2 or 8 bytes) the uniform-degree bid scans read each value with on the resident CSR.
int sla_scan_value_bytes(const sla_ctx* ctx, uint32_t* value_bytes) {
    if (!ctx || !value_bytes) return SLA_ERR_INVALID;
    *value_bytes = (ctx->has_csr && narrow_scan_ptr(ctx) && !use_stream_scan(ctx)) ? 2u : 8u;
    return SLA_OK;
}

int sla_upload_csr_device(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, const uint32_t* d_row_ptr,
                          const uint32_t* d_column_indices, const double* d_values, uint64_t nnz) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!d_row_ptr || !d_column_indices || !d_values) return fail(ctx, SLA_ERR_INVALID, "null input array");
    int rc = check_shape(ctx, num_rows, num_cols, nnz);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure_capacity(ctx, num_rows, num_cols, nnz))) return rc;
    CU(cudaMemcpyAsync(ctx->d_row_ptr, d_row_ptr, ((size_t)num_rows + 1) * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_cols, d_column_indices, (size_t)nnz * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_vals, d_values, (size_t)nnz * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return finish_csr(ctx, num_rows, num_cols, nnz, true);
}

static int make_spec(sla_ctx* ctx, sla_synth::Spec* s, uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed,
                     uint32_t value_lo, uint32_t value_hi, int planted) {
    if (sla_hostgen::make_spec(s, num_rows, num_cols, k, seed, value_lo, value_hi, planted, 0))
        return fail(ctx, SLA_ERR_INVALID, "generator arguments: k in [1, num_cols], value_hi > value_lo, num_rows * k < u32::MAX");
    return SLA_OK;
}

int sla_generate_device(sla_ctx* ctx, uint32_t num_rows, uint32_t num_cols, uint32_t k, uint64_t seed,
                        uint32_t value_lo, uint32_t value_hi, int planted) {
    if (!ctx) return SLA_ERR_INVALID;
    const uint64_t nnz = (uint64_t)num_rows * k;
    int rc = check_shape(ctx, num_rows, num_cols, nnz);
    if (rc) return rc;
    sla_synth::Spec s;
    if ((rc = make_spec(ctx, &s, num_rows, num_cols, k, seed, value_lo, value_hi, planted))) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure_capacity(ctx, num_rows, num_cols, nnz))) return rc;
    generate_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(s, 0u, num_rows, ctx->d_row_ptr, ctx->d_cols, ctx->d_vals);
    return finish_csr(ctx, num_rows, num_cols, nnz, true);
}

int sla_generate_device_shard(sla_ctx* ctx, uint32_t global_rows, uint32_t num_cols, uint32_t k, uint64_t seed,
                              uint32_t value_lo, uint32_t value_hi, int planted, uint32_t row_begin, uint32_t row_count) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!row_count || (uint64_t)row_begin + row_count > global_rows) return fail(ctx, SLA_ERR_INVALID, "bad shard range");
    const uint64_t nnz = (uint64_t)row_count * k;
    int rc = check_shape(ctx, row_count, num_cols, nnz);
    if (rc) return rc;
    if (global_rows > num_cols) return fail(ctx, SLA_ERR_INVALID, "num_rows must be <= num_cols");
    sla_synth::Spec s;
    if ((rc = make_spec(ctx, &s, global_rows, num_cols, k, seed, value_lo, value_hi, planted))) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = ensure_capacity(ctx, row_count, num_cols, nnz))) return rc;
    generate_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(s, row_begin, row_count, ctx->d_row_ptr, ctx->d_cols,
                                                                     ctx->d_vals);
    return finish_csr(ctx, row_count, num_cols, nnz, true);
}

int sla_khosla_solve(sla_ctx* ctx, int maximize, double eps, uint32_t* person_to_object, uint32_t* object_to_person,
                     double* prices, sla_stats* stats) {
    return solve_common(ctx, SLA_ALGO_KHOSLA, maximize, eps, NAN, 0, person_to_object, object_to_person, prices, stats);
}

int sla_forward_solve(sla_ctx* ctx, int maximize, double eps, double start_eps, uint32_t max_iterations,
                      uint32_t* person_to_object, uint32_t* object_to_person, double* prices, sla_stats* stats) {
    return solve_common(ctx, SLA_ALGO_FORWARD, maximize, eps, start_eps, max_iterations, person_to_object,
                        object_to_person, prices, stats);
}

int sla_download_solution(sla_ctx* ctx, uint32_t* person_to_object, uint32_t* object_to_person, double* prices) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->has_solution) return fail(ctx, SLA_ERR_STATE, "no resident solution");
    NvtxRange nvtx_dl("sla_download_solution");
    CU(cudaSetDevice(ctx->device));
    if (person_to_object)
        CU(cudaMemcpyAsync(person_to_object, ctx->d_p2o, (size_t)ctx->n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (object_to_person)
        CU(cudaMemcpyAsync(object_to_person, ctx->d_o2p, (size_t)ctx->n_cols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (prices) CU(cudaMemcpyAsync(prices, ctx->d_prices, (size_t)ctx->n_cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return SLA_OK;
}

int sla_get_objective(sla_ctx* ctx, double* objective) {
    if (!ctx || !objective) return SLA_ERR_INVALID;
    if (!ctx->has_solution) return fail(ctx, SLA_ERR_STATE, "no resident solution");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    const uint32_t flip = ctx->dev_sign < 0 ? 0x80000000u : 0u;
    objective_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->n_rows, flip, ctx->d_partial);
    CU(cudaMemcpyAsync(ctx->h_partial, ctx->d_partial, (size_t)ctx->grid_wide * sizeof(double), cudaMemcpyDeviceToHost,
                       ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    double sum = 0.0;
    for (int b = 0; b < ctx->grid_wide; ++b) sum += ctx->h_partial[b];
    // reference src/solver.rs:111-137: the sign is taken from the (current, possibly negated) first value
    const double first_eff = ctx->dev_sign > 0 ? ctx->first_value : -ctx->first_value;
    *objective = (first_eff >= 0.0) ? sum : -sum;
    return SLA_OK;
}

int sla_ecs_satisfied(sla_ctx* ctx, double eps, double toleration, int* satisfied) {
    if (!ctx || !satisfied) return SLA_ERR_INVALID;
    if (!ctx->has_solution) return fail(ctx, SLA_ERR_STATE, "no resident solution");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    const uint32_t flip = ctx->dev_sign < 0 ? 0x80000000u : 0u;
    CU(cudaMemsetAsync(ctx->d_scratch, 0, 16 * sizeof(uint32_t), ctx->stream));
    ecs_check_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->n_rows, ctx->n_cols, flip, eps, toleration,
                                                                      ctx->d_scratch);
    CU(cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    *satisfied = ctx->h_scratch[0] ? 0 : 1;
    return SLA_OK;
}

int sla_validate_matching(sla_ctx* ctx, uint32_t* num_unassigned, int* consistent) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->has_solution) return fail(ctx, SLA_ERR_STATE, "no resident solution");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    CU(cudaMemsetAsync(ctx->d_scratch, 0, 16 * sizeof(uint32_t), ctx->stream));
    validate_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->n_rows, ctx->n_cols, ctx->d_scratch);
    CU(cudaMemcpyAsync(ctx->h_scratch, ctx->d_scratch, 16 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (num_unassigned) *num_unassigned = ctx->h_scratch[0];
    if (consistent) *consistent = ctx->h_scratch[1] ? 0 : 1;
    return SLA_OK;
}

// ---- host-memory helpers for the wrappers (pinned staging keeps the H2D / D2H copies at PCIe speed) ----
int sla_host_alloc(size_t bytes, void** out) {
    if (!out) return SLA_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        g_create_error = std::string("cudaMallocHost: ") + cudaGetErrorString(e);
        *out = nullptr;
        return SLA_ERR_ALLOC;
    }
    return SLA_OK;
}

void sla_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

// In-place negation of the host copy of `values` (the observable half of solver.rs:214-216), split over threads.
void sla_host_negate_f64(double* values, size_t n, int threads) {
    if (!values || !n) return;
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    if (n < (size_t)1 << 16) threads = 1;
    auto work = [values](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) values[i] = -values[i];
    };
    std::vector<std::thread> pool;
    const size_t chunk = (n + (size_t)threads - 1) / (size_t)threads;
    for (int t = 1; t < threads; ++t) {
        const size_t lo = (size_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
        if (lo < hi) pool.emplace_back(work, lo, hi);
    }
    work(0, chunk < n ? chunk : n);
    for (auto& th : pool) th.join();
}

// Development aid (not part of include/sla.h): raw tail-engine cycle counters of the last solve.
int sla_debug_counters(sla_ctx* ctx, uint64_t* out24) {
    if (!ctx || !out24) return SLA_ERR_INVALID;
    for (int i = 0; i < 24; ++i) out24[i] = ctx->h_state->dbg[i];
    return SLA_OK;
}

int sla_get_round_profile(sla_ctx* ctx, sla_round_profile* out, size_t capacity, size_t* count) {
    if (!ctx || !count) return SLA_ERR_INVALID;
    *count = ctx->profile.size();
    if (out) {
        const size_t n = ctx->profile.size() < capacity ? ctx->profile.size() : capacity;
        if (n) memcpy(out, ctx->profile.data(), n * sizeof(sla_round_profile));
    }
    return SLA_OK;
}

}  // extern "C"

#include "sla_batch.cuh"
#include "sla_part.cuh"
#include "sla_mesh.cuh"
