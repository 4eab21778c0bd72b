// sla_batch.cuh -- batch of independent instances, one CTA per instance (BASELINE.json config 4).
//
// The reference has no batch API (its bench harness clones one solver per problem, benches/benchmark.rs:109,137);
// semantically every instance is its own KhoslaSolver / ForwardAuctionSolver::solve: own sign normalisation
// (solver.rs:207-216), own eps / threshold / toleration, own eps-scaling phases.  The whole auction of an instance
// runs inside one CTA with all of its state (prices, owners, assignment, queue, bid words) in shared memory; only
// the CSR rows are read from global memory (L1/L2 resident).  Rounds are the same synchronous Jacobi rounds as in the
// single-instance engines, so every instance equals oracle/jacobi_model.c bit for bit.
#pragma once

namespace sla {

// 128 threads and >= 8 CTAs per SM measured best on cfg4 (profiles/README.md): every CTA is a latency-bound chain of
// rounds, so the SM needs many co-resident instances to keep its issue slots busy.
#ifndef SLA_BATCH_THREADS
#define SLA_BATCH_THREADS 128
#endif
#ifndef SLA_BATCH_MINBLOCKS
#define SLA_BATCH_MINBLOCKS 8
#endif
constexpr int kBatchThreads = SLA_BATCH_THREADS;

struct DevInstStats {   // per-instance result scalars (expanded into sla_stats on the host)
    uint32_t num_unassigned, nits, nreductions, optimal;
    double eps;
    unsigned long long rounds, bids, bid_arcs;
    uint32_t dropped, values_negated;
    uint32_t restarts, pad_;   // restarts: 1 when the Khosla eps-schedule was abandoned for the plain rounds
};

struct BatchParams {
    const uint32_t* __restrict__ row_off;   // n_inst + 1
    const uint32_t* __restrict__ col_off;   // n_inst + 1
    const uint32_t* __restrict__ row_ptr;   // total_rows + 1, global arc offsets
    const uint32_t* __restrict__ cols;      // instance-local column indices
    const double* __restrict__ vals;
    uint32_t* p2o;                          // total_rows (instance-local object index)
    uint32_t* o2p;                          // total_cols (instance-local person index)
    double* prices;                         // total_cols
    DevInstStats* stats;                    // n_inst
    uint32_t n_inst, max_rows, max_cols;
    uint32_t algo, maximize, max_iterations;
    uint32_t khosla_scaling;                // Khosla rounds under an eps-schedule on square instances (finish_if_possible)
    double eps_in, start_eps_in;            // NaN = None
};

__device__ __forceinline__ double batch_toleration(double c) {
    // reference src/solver.rs:144-146
    const double l = log2(c + 1e-7);
    const uint32_t li = !(l > 0.0) ? 0u : (l >= 4294967295.0 ? 4294967295u : (uint32_t)l);
    const uint32_t e = 53u - li;
    const unsigned long long pw = e < 64 ? (1ull << e) : 0ull;
    return 1.0 / (double)pw;
}

template <int LPR>
__global__ void __launch_bounds__(kBatchThreads, SLA_BATCH_MINBLOCKS) batch_kernel(const BatchParams p) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    // ---- shared-memory carve-up (sizes from the largest instance of the batch) ----
    double* s_prices = reinterpret_cast<double*>(s_dyn);                                   // max_cols
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(s_prices + p.max_cols);   // max_cols
    double* s_bid = reinterpret_cast<double*>(s_best + p.max_cols);                        // max_rows
    uint32_t* s_o2p = reinterpret_cast<uint32_t*>(s_bid + p.max_rows);                     // max_cols
    uint32_t* s_p2o = s_o2p + p.max_cols;                                                  // max_rows
    uint32_t* s_q0 = s_p2o + p.max_rows;                                                   // max_rows
    uint32_t* s_q1 = s_q0 + p.max_rows;                                                    // max_rows
    uint32_t* s_obj = s_q1 + p.max_rows;                                                   // max_rows
    uint32_t* s_prev = s_obj + p.max_rows;                                                 // max_rows

    __shared__ unsigned long long s_word[32];
    __shared__ uint32_t s_warp_cnt[kBatchThreads / 32];
    __shared__ unsigned long long s_red_min, s_red_max;
    __shared__ uint32_t s_next_len, s_flag;
    __shared__ unsigned long long s_arcs;
    __shared__ uint32_t s_dropped;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane32 = tid & 31;
    constexpr int NGROUPS = kBatchThreads / LPR;
    const int lane = tid % LPR;
    const uint32_t group = tid / LPR;

    for (uint32_t inst = blockIdx.x; inst < p.n_inst; inst += gridDim.x) {
        const uint32_t r0 = p.row_off[inst], N = p.row_off[inst + 1] - r0;
        const uint32_t c0 = p.col_off[inst], M = p.col_off[inst + 1] - c0;
        const uint32_t* row_ptr = p.row_ptr + r0;
        const uint32_t arc_lo = row_ptr[0], arc_hi = row_ptr[N];

        // ---- prologue: value range of this instance + sign normalisation (solver.rs:207-216, ksparse.rs:171-179) ----
        if (tid == 0) { s_red_min = ~0ull; s_red_max = 0ull; s_arcs = 0; s_dropped = 0; }
        __syncthreads();
        {
            unsigned long long kmin = ~0ull, kmax = 0ull;
            for (uint32_t g = arc_lo + tid; g < arc_hi; g += kBatchThreads) {
                const unsigned long long k = f64_order_key(p.vals[g]);
                kmin = k < kmin ? k : kmin;
                kmax = k > kmax ? k : kmax;
            }
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) {
                const unsigned long long omin = __shfl_xor_sync(0xffffffffu, kmin, m);
                const unsigned long long omax = __shfl_xor_sync(0xffffffffu, kmax, m);
                kmin = omin < kmin ? omin : kmin;
                kmax = omax > kmax ? omax : kmax;
            }
            if (lane32 == 0) { atomicMin(&s_red_min, kmin); atomicMax(&s_red_max, kmax); }
        }
        for (uint32_t j = tid; j < M; j += kBatchThreads) { s_prices[j] = 0.0; s_o2p[j] = SLA_DEV_NONE; s_best[j] = 0ull; }
        for (uint32_t i = tid; i < N; i += kBatchThreads) { s_p2o[i] = SLA_DEV_NONE; s_q0[i] = i; }
        __syncthreads();

        const double v_min = order_key_to_f64_host(s_red_min), v_max = order_key_to_f64_host(s_red_max);
        const double first = (arc_hi > arc_lo) ? p.vals[arc_lo] : 0.0;
        const bool negate = (p.maximize != 0) != (first >= 0.0);
        const uint32_t sign_flip = negate ? 0x80000000u : 0u;
        const double w_min = negate ? -v_max : v_min, w_max = negate ? -v_min : v_max;
        const uint32_t algo = p.algo;
        const uint32_t pbits = 32u - (uint32_t)__clz((int)(N > 1 ? N - 1 : 1));
        double eps, threshold = 0.0, target = 0.0, tol = 0.0;
        uint32_t start_opt = 0;
        const uint32_t max_it = p.max_iterations ? p.max_iterations : 100000u;
        if (algo == ALGO_KHOSLA) {
            const double m = (double)M;                                            // ksparse.rs:160-181
            eps = isnan(p.eps_in) ? 1.0 / m : p.eps_in;
            threshold = (m / 2.0) * (w_max - w_min + eps);
            target = eps;
        } else {
            target = isnan(p.eps_in) ? 1.0 / (double)N : p.eps_in;                 // symmetric.rs:229-273
            const double c = fmax(fabs(w_min), fabs(w_max));
            tol = batch_toleration(c);
            start_opt = !isnan(p.start_eps_in) ? (p.start_eps_in < target ? 1u : 0u) : 0u;
            if (N != M) { start_opt = 1u; eps = target - 2.220446049250313e-16; }
            else eps = !isnan(p.start_eps_in) ? p.start_eps_in : c / 2.0;
        }

        // Khosla on a square instance: rounds under the eps-schedule c/2, x0.15, ..., caller's eps; a phase that drops
        // anybody makes the instance start over with the plain rounds (same rule as finish_if_possible)
        bool kscale = false;
        uint32_t restarts = 0;
        if (algo == ALGO_KHOSLA && p.khosla_scaling && N == M) {
            const double c = fmax(fabs(w_min), fabs(w_max));
            if (c / 2.0 > eps) { kscale = true; eps = c / 2.0; }
        }
        uint32_t qlen = N, nits = 0, nreductions = 0, optimal = 0;
        unsigned long long rounds = 0, bids = 0, my_arcs = 0;
        uint32_t my_dropped = 0;
        uint32_t* sq = s_q0;
        uint32_t* nq = s_q1;
        bool zero = true;   // all prices exactly zero: first round only

        while (true) {
            const bool small = qlen <= 32u;
            // ---- bidding phase ----
            if (small) {
                // one warp per bidder, several passes when there are more bidders than warps (<= 32 bidders)
                for (uint32_t q = (uint32_t)warp; q < qlen; q += kBatchThreads / 32) {
                    const uint32_t i = sq[q];
                    const uint32_t a = __ldg(row_ptr + i), b = __ldg(row_ptr + i + 1);
                    WarpChoice c;
                    if (zero) c = warp_bid_scan<PRICE_ZERO, OWN_NONE>(p.cols, p.vals, s_prices, s_o2p, a, b, sign_flip, lane32);
                    else      c = warp_bid_scan<PRICE_SMEM, OWN_SMEM>(p.cols, p.vals, s_prices, s_o2p, a, b, sign_flip, lane32);
                    if (lane32 == 0) {
                        uint32_t owner;
                        const Bid r = zero ? make_bid_warp<PRICE_ZERO, OWN_NONE>(c, algo, eps, threshold, s_prices, s_o2p, &owner)
                                           : make_bid_warp<PRICE_SMEM, OWN_SMEM>(c, algo, eps, threshold, s_prices, s_o2p, &owner);
                        my_arcs += (unsigned long long)(b - a);
                        if (r.dropped) {
                            s_obj[q] = SLA_DEV_NONE;
                            my_dropped += 1;
                        } else {
                            s_obj[q] = r.obj;
                            s_bid[q] = r.bid;
                            s_prev[q] = owner;
                            s_word[q] = (r.bid == r.bid) ? pack_bid(r.bid, i, pbits) : 0ull;
                        }
                    }
                }
            } else {
            for (uint32_t base = 0; base < qlen; base += NGROUPS) {
                if (base + (uint32_t)(warp * 32) / LPR >= qlen) break;   // warp-uniform: idle warps leave
                const uint32_t q = base + group;
                const bool valid = q < qlen;
                uint32_t i = 0, a = 0, b = 0;
                if (valid) { i = sq[q]; a = __ldg(row_ptr + i); b = __ldg(row_ptr + i + 1); }
                Choice c;
                choice_init(c);
                if (zero) scan_row<LPR, PRICE_ZERO, false>(c, p.cols, p.vals, s_prices, a, b, sign_flip, lane);
                else      scan_row<LPR, PRICE_SMEM, false>(c, p.cols, p.vals, s_prices, a, b, sign_flip, lane);
                choice_group_reduce<LPR>(c);
                if (valid && lane == 0) {
                    const Bid r = zero ? make_bid<PRICE_ZERO>(c, algo, eps, threshold, s_prices)
                                       : make_bid<PRICE_SMEM>(c, algo, eps, threshold, s_prices);
                    my_arcs += (unsigned long long)(b - a);
                    if (r.dropped) {
                        s_obj[q] = SLA_DEV_NONE;
                        my_dropped += 1;
                    } else {
                        s_obj[q] = r.obj;
                        s_bid[q] = r.bid;
                        s_prev[q] = s_o2p[r.obj];
                        if (r.bid == r.bid) atomicMax(s_best + r.obj, pack_bid(r.bid, i, pbits));
                    }
                }
            }
            }
            __syncthreads();

            // ---- assignment phase + compaction ----
            uint32_t out = 0;
            if (small) {
                if (warp == 0) {
                    const uint32_t q = (uint32_t)lane32;
                    const bool live = q < qlen;
                    const uint32_t j = live ? s_obj[q] : SLA_DEV_NONE;
                    const bool bidding = j != SLA_DEV_NONE;
                    const uint32_t i = live ? sq[q] : 0u;
                    const unsigned long long w = bidding ? s_word[q] : 0ull;
                    const double bid = bidding ? s_bid[q] : 0.0;
                    const uint32_t prev = bidding ? s_prev[q] : SLA_DEV_NONE;
// Conflict detection among the <= 32 bidders of a small round.  MATCH.ANY walks the distinct keys one by one, so the idle
// lanes share one key (variant 1: cfg4 42.5 -> 40.5 ms against unique keys per idle lane, variant 0; an all-pairs
// shuffle loop, variant 2, is no faster).
#ifndef SLA_BATCH_RESOLVE
#define SLA_BATCH_RESOLVE 1
#endif
                    bool won = bidding && (w != 0ull);
#if SLA_BATCH_RESOLVE == 2
                    // all-pairs comparison over the (few) bidders of the round: two shuffles per bidder, all independent
                    for (uint32_t r = 0; r < qlen; ++r) {
                        const uint32_t oj = __shfl_sync(0xffffffffu, j, (int)r);
                        const unsigned long long ow = __shfl_sync(0xffffffffu, w, (int)r);
                        won = won && !(oj == j && ow > w);
                    }
#else
#if SLA_BATCH_RESOLVE == 1
                    const uint32_t peers = __match_any_sync(0xffffffffu, j);   // idle lanes share the key SLA_DEV_NONE
#else
                    const uint32_t peers = __match_any_sync(0xffffffffu, bidding ? j : (0xFFFFFF00u + (uint32_t)lane32));
#endif
                    if (__any_sync(0xffffffffu, bidding && (peers & (peers - 1u)) != 0u)) {
                        for (uint32_t r = 0; r < qlen; ++r) {
                            const unsigned long long ow = __shfl_sync(0xffffffffu, w, (int)r);
                            won = won && !(((peers >> r) & 1u) && ow > w);
                        }
                    }
#endif
                    uint32_t emit = SLA_DEV_NONE;
                    if (bidding) {
                        if (won) {
                            s_prices[j] = bid;
                            s_o2p[j] = i;
                            s_p2o[i] = j;
                            if (prev != SLA_DEV_NONE) { s_p2o[prev] = SLA_DEV_NONE; emit = prev; }
                        } else {
                            emit = i;
                        }
                    }
                    const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
                    if (emit != SLA_DEV_NONE) nq[__popc(ballot & ((1u << lane32) - 1u))] = emit;
                    if (lane32 == 0) s_next_len = __popc(ballot);
                }
                __syncthreads();
                out = s_next_len;
            } else {
                for (uint32_t base = 0; base < qlen; base += kBatchThreads) {
                    const uint32_t q = base + tid;
                    uint32_t emit = SLA_DEV_NONE;
                    if (q < qlen) {
                        const uint32_t j = s_obj[q];
                        if (j != SLA_DEV_NONE) {
                            const uint32_t i = sq[q];
                            const double bid = s_bid[q];
                            const bool won = (bid == bid) && (s_best[j] == pack_bid(bid, i, pbits));
                            if (won) {
                                const uint32_t prev = s_prev[q];
                                s_prices[j] = bid;
                                s_o2p[j] = i;
                                s_p2o[i] = j;
                                if (prev != SLA_DEV_NONE) { s_p2o[prev] = SLA_DEV_NONE; emit = prev; }
                            } else {
                                emit = i;
                            }
                        }
                    }
                    const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
                    if (lane32 == 0) s_warp_cnt[warp] = __popc(ballot);
                    __syncthreads();
                    uint32_t off = 0, total = 0;
#pragma unroll
                    for (int w = 0; w < kBatchThreads / 32; ++w) {
                        const uint32_t cnt = s_warp_cnt[w];
                        off += (w < warp) ? cnt : 0u;
                        total += cnt;
                    }
                    if (emit != SLA_DEV_NONE) nq[out + off + __popc(ballot & ((1u << lane32) - 1u))] = emit;
                    out += total;
                    __syncthreads();
                }
                // all words have been compared: clear the ones this round touched (every slot's object, winners or not)
                for (uint32_t q = tid; q < qlen; q += kBatchThreads) {
                    const uint32_t j = s_obj[q];
                    if (j != SLA_DEV_NONE) s_best[j] = 0ull;
                }
                __syncthreads();
            }

            bids += qlen;
            rounds += 1;
            qlen = out;
            { uint32_t* t = sq; sq = nq; nq = t; }
            zero = false;
            if (algo == ALGO_KHOSLA) {
                if (qlen != 0) continue;
                if (!kscale) break;
                // end of a phase of the eps-schedule: did anybody hit the price threshold?
                if (tid == 0) s_flag = 0;
                __syncthreads();
                if (my_dropped) s_flag = 1;
                __syncthreads();
                const bool any_dropped = s_flag != 0;
                __syncthreads();
                if (any_dropped) {
                    kscale = false;
                    restarts = 1;
                    eps = target;
                    my_dropped = 0;
                    for (uint32_t j = tid; j < M; j += kBatchThreads) s_prices[j] = 0.0;
                } else if (eps > target) {
                    eps *= 0.15;
                    if (eps < target) eps = target;
                    nreductions += 1;
                } else {
                    break;
                }
                for (uint32_t i = tid; i < N; i += kBatchThreads) { s_p2o[i] = SLA_DEV_NONE; sq[i] = i; }
                for (uint32_t j = tid; j < M; j += kBatchThreads) s_o2p[j] = SLA_DEV_NONE;
                qlen = N;
                __syncthreads();
                continue;
            }
            // ---- Forward: eps-scaling control (symmetric.rs:275-329) ----
            nits += 1;
            if (qlen == 0) {
                bool is_optimal = start_opt != 0;
                if (!is_optimal) {
                    // eps-CS check against target_eps (solver.rs:154-189), all threads, state in shared memory
                    if (tid == 0) s_flag = 0;
                    __syncthreads();
                    bool violated = false;
                    for (uint32_t base = 0; base < N; base += NGROUPS) {
                        const uint32_t i = base + group;
                        const bool valid = i < N;
                        uint32_t a = 0, b = 0, j = 0;
                        if (valid) { a = __ldg(row_ptr + i); b = __ldg(row_ptr + i + 1); j = s_p2o[i]; }
                        uint32_t cpos = 0;
                        double cval = neg_inf();
                        for (uint32_t g = a + lane; g < b; g += LPR)
                            if (__ldg(p.cols + g) == j) { cpos = g + 1u; cval = __ldg(p.vals + g); }
#pragma unroll
                        for (int m = LPR / 2; m >= 1; m >>= 1) {
                            const uint32_t op = __shfl_xor_sync(0xffffffffu, cpos, m);
                            const double ov = __shfl_xor_sync(0xffffffffu, cval, m);
                            if (op > cpos) { cpos = op; cval = ov; }
                        }
                        if (valid) {
                            if (j >= M) {
                                violated = true;
                            } else {
                                const double chosen = (cpos == 0u) ? neg_inf()
                                    : __hiloint2double(__double2hiint(cval) ^ (int)sign_flip, __double2loint(cval));
                                const double lhs = chosen - s_prices[j] + tol;
                                for (uint32_t g = a + lane; g < b; g += LPR) {
                                    const double raw = __ldg(p.vals + g);
                                    const double v = __hiloint2double(__double2hiint(raw) ^ (int)sign_flip, __double2loint(raw));
                                    if (lhs < v - s_prices[__ldg(p.cols + g)] - target) violated = true;
                                }
                            }
                        }
                    }
                    if (violated) s_flag = 1;
                    __syncthreads();
                    is_optimal = (s_flag == 0);
                    __syncthreads();
                }
                if (is_optimal) { optimal = 1; break; }
                if (eps < target) break;
                eps *= 0.15;
                for (uint32_t i = tid; i < N; i += kBatchThreads) { s_p2o[i] = SLA_DEV_NONE; sq[i] = i; }
                for (uint32_t j = tid; j < M; j += kBatchThreads) s_o2p[j] = SLA_DEV_NONE;
                qlen = N;
                nreductions += 1;
                __syncthreads();
            }
            if (nits >= max_it) break;
        }

        // ---- epilogue: results to global memory ----
        if (my_arcs) atomicAdd(&s_arcs, my_arcs);
        if (my_dropped) atomicAdd(&s_dropped, my_dropped);
        __syncthreads();
        for (uint32_t i = tid; i < N; i += kBatchThreads) p.p2o[r0 + i] = s_p2o[i];
        for (uint32_t j = tid; j < M; j += kBatchThreads) { p.o2p[c0 + j] = s_o2p[j]; p.prices[c0 + j] = s_prices[j]; }
        if (tid == 0) {
            DevInstStats st;
            st.nreductions = nreductions;
            st.optimal = optimal;
            st.eps = eps;
            st.rounds = rounds;
            st.bids = bids;
            st.bid_arcs = s_arcs;
            st.dropped = s_dropped;
            st.values_negated = negate ? 1u : 0u;
            st.restarts = restarts;
            st.pad_ = 0u;
            if (algo == ALGO_KHOSLA) { st.nits = (uint32_t)bids; st.num_unassigned = s_dropped; }
            else { st.nits = nits; st.num_unassigned = qlen; }
            p.stats[inst] = st;
        }
        __syncthreads();
    }
}

// Batch generator: instance b (global id first_id + b) is sla_generate_* with seed `seed + first_id + b`.
__global__ void __launch_bounds__(kWideThreads) batch_generate_kernel(sla_synth::Spec tmpl, const uint32_t n_inst,
                                                                      const uint32_t first_id, uint32_t* __restrict__ row_off,
                                                                      uint32_t* __restrict__ col_off, uint32_t* __restrict__ row_ptr,
                                                                      uint32_t* __restrict__ cols, double* __restrict__ vals) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint32_t N = tmpl.num_rows, M = tmpl.num_cols, K = tmpl.k;
    const uint32_t total = n_inst * N;
    for (uint32_t r = tid; r < total; r += stride) {
        const uint32_t b = r / N, i = r - b * N;
        sla_synth::Spec s = tmpl;
        s.seed = tmpl.seed + first_id + b;
        sla_synth::finish_spec(s);
        const size_t off = (size_t)r * K;
        sla_synth::make_row(s, i, cols + off, vals + off);
        row_ptr[r] = (uint32_t)off;
    }
    for (uint32_t b = tid; b <= n_inst; b += stride) { row_off[b] = b * N; col_off[b] = b * M; }
    if (tid == 0) row_ptr[total] = total * K;
}

}  // namespace sla

struct sla_batch_state {
    uint32_t n_inst = 0, total_rows = 0, total_cols = 0, max_rows = 0, max_cols = 0;
    uint64_t nnz = 0;
    int lpr = 4;
    size_t cap_inst = 0, cap_rows = 0, cap_cols = 0, cap_arcs = 0;
    uint32_t *d_row_off = nullptr, *d_col_off = nullptr, *d_row_ptr = nullptr, *d_cols = nullptr;
    double* d_vals = nullptr;
    uint32_t *d_p2o = nullptr, *d_o2p = nullptr;
    double* d_prices = nullptr;
    sla::DevInstStats* d_stats = nullptr;
    std::vector<sla::DevInstStats> h_stats;
    bool ready = false;
};

namespace {

int batch_reserve(sla_ctx* ctx, size_t n_inst, size_t rows, size_t cols, size_t arcs) {
    if (!ctx->batch) ctx->batch = new sla_batch_state();
    sla_batch_state* b = ctx->batch;
    int rc;
    if (n_inst > b->cap_inst) {
        if ((rc = dev_alloc(ctx, &b->d_row_off, n_inst + 1)) || (rc = dev_alloc(ctx, &b->d_col_off, n_inst + 1)) ||
            (rc = dev_alloc(ctx, &b->d_stats, n_inst)))
            return rc;
        b->cap_inst = n_inst;
    }
    if (rows > b->cap_rows) {
        if ((rc = dev_alloc(ctx, &b->d_row_ptr, rows + 8)) || (rc = dev_alloc(ctx, &b->d_p2o, rows))) return rc;
        b->cap_rows = rows;
    }
    if (cols > b->cap_cols) {
        if ((rc = dev_alloc(ctx, &b->d_o2p, cols)) || (rc = dev_alloc(ctx, &b->d_prices, cols))) return rc;
        b->cap_cols = cols;
    }
    if (arcs > b->cap_arcs) {
        if ((rc = dev_alloc(ctx, &b->d_cols, arcs + 8)) || (rc = dev_alloc(ctx, &b->d_vals, arcs + 8))) return rc;
        CU(cudaMemsetAsync(b->d_cols, 0, (arcs + 8) * sizeof(uint32_t), ctx->stream));
        CU(cudaMemsetAsync(b->d_vals, 0, (arcs + 8) * sizeof(double), ctx->stream));
        b->cap_arcs = arcs;
    }
    return SLA_OK;
}

size_t batch_smem_bytes(uint32_t max_rows, uint32_t max_cols) {
    return (size_t)max_cols * (8 + 8 + 4) + (size_t)max_rows * (8 + 4 * 5);
}

template <int LPR>
int batch_launch_t(sla_ctx* ctx, const sla::BatchParams& bp, size_t smem, int grid) {
    CU(cudaFuncSetAttribute(sla::batch_kernel<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sla::batch_kernel<LPR><<<grid, sla::kBatchThreads, smem, ctx->stream>>>(bp);
    return SLA_OK;
}

template <int LPR>
int batch_occupancy_t(size_t smem) {
    int occ = 0;
    cudaFuncSetAttribute(sla::batch_kernel<LPR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sla::batch_kernel<LPR>, sla::kBatchThreads, smem);
    return occ < 1 ? 1 : occ;
}

}  // namespace

extern "C" {

void sla_batch_free(sla_ctx* ctx) {
    if (!ctx || !ctx->batch) return;
    sla_batch_state* b = ctx->batch;
    cudaFree(b->d_row_off); cudaFree(b->d_col_off); cudaFree(b->d_row_ptr); cudaFree(b->d_cols); cudaFree(b->d_vals);
    cudaFree(b->d_p2o); cudaFree(b->d_o2p); cudaFree(b->d_prices); cudaFree(b->d_stats);
    delete b;
    ctx->batch = nullptr;
}

int sla_batch_upload(sla_ctx* ctx, uint32_t num_instances, const uint32_t* row_off, const uint32_t* col_off,
                     const uint32_t* row_ptr, const uint32_t* column_indices, const double* values) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!num_instances || !row_off || !col_off || !row_ptr || !column_indices || !values)
        return fail(ctx, SLA_ERR_INVALID, "null or empty batch input");
    CU(cudaSetDevice(ctx->device));
    const uint32_t total_rows = row_off[num_instances], total_cols = col_off[num_instances];
    const uint64_t nnz = row_ptr[total_rows];
    uint32_t max_rows = 0, max_cols = 0;
    for (uint32_t b = 0; b < num_instances; ++b) {
        if (row_off[b + 1] <= row_off[b] || col_off[b + 1] <= col_off[b])
            return fail(ctx, SLA_ERR_INVALID, "every instance needs at least one row and one column");
        const uint32_t n = row_off[b + 1] - row_off[b], m = col_off[b + 1] - col_off[b];
        if (n > m) return fail(ctx, SLA_ERR_INVALID, "num_rows must be <= num_cols in every instance");   // solver.rs:192
        if (row_ptr[row_off[b + 1]] <= row_ptr[row_off[b]]) return fail(ctx, SLA_ERR_INVALID, "instance without arcs");
        max_rows = n > max_rows ? n : max_rows;
        max_cols = m > max_cols ? m : max_cols;
    }
    if (batch_smem_bytes(max_rows, max_cols) > 200 * 1024)
        return fail(ctx, SLA_ERR_INVALID, "instance too large for the one-CTA-per-instance batch engine (use the single-instance API)");
    int rc = batch_reserve(ctx, num_instances, total_rows, total_cols, nnz);
    if (rc) return rc;
    sla_batch_state* b = ctx->batch;
    b->ready = false;
    CU(cudaMemcpyAsync(b->d_row_off, row_off, ((size_t)num_instances + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(b->d_col_off, col_off, ((size_t)num_instances + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(b->d_row_ptr, row_ptr, ((size_t)total_rows + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(b->d_cols, column_indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(b->d_vals, values, (size_t)nnz * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    b->n_inst = num_instances; b->total_rows = total_rows; b->total_cols = total_cols; b->nnz = nnz;
    b->max_rows = max_rows; b->max_cols = max_cols;
    b->lpr = pick_lpr(nnz, total_rows);
    b->ready = true;
    return SLA_OK;
}

int sla_batch_generate_device(sla_ctx* ctx, uint32_t num_instances, uint32_t first_instance_id, uint32_t num_rows,
                              uint32_t num_cols, uint32_t k, uint64_t seed, uint32_t value_lo, uint32_t value_hi,
                              int planted) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!num_instances) return fail(ctx, SLA_ERR_INVALID, "empty batch");
    if (num_rows > num_cols) return fail(ctx, SLA_ERR_INVALID, "num_rows must be <= num_cols");
    sla_synth::Spec s;
    int rc = make_spec(ctx, &s, num_rows, num_cols, k, seed, value_lo, value_hi, planted);
    if (rc) return rc;
    const uint64_t total_rows = (uint64_t)num_instances * num_rows, total_cols = (uint64_t)num_instances * num_cols;
    const uint64_t nnz = total_rows * k;
    if (nnz >= 0xFFFFFFFFull || total_cols >= 0xFFFFFFFFull) return fail(ctx, SLA_ERR_INVALID, "batch exceeds u32 offsets");
    if (batch_smem_bytes(num_rows, num_cols) > 200 * 1024)
        return fail(ctx, SLA_ERR_INVALID, "instance too large for the one-CTA-per-instance batch engine");
    CU(cudaSetDevice(ctx->device));
    if ((rc = batch_reserve(ctx, num_instances, total_rows, total_cols, nnz))) return rc;
    sla_batch_state* b = ctx->batch;
    sla::batch_generate_kernel<<<ctx->grid_wide, sla::kWideThreads, 0, ctx->stream>>>(s, num_instances, first_instance_id, b->d_row_off,
                                                                                  b->d_col_off, b->d_row_ptr, b->d_cols, b->d_vals);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    b->n_inst = num_instances; b->total_rows = (uint32_t)total_rows; b->total_cols = (uint32_t)total_cols; b->nnz = nnz;
    b->max_rows = num_rows; b->max_cols = num_cols;
    b->lpr = pick_lpr(nnz, (uint32_t)total_rows);
    b->ready = true;
    return SLA_OK;
}

int sla_batch_solve(sla_ctx* ctx, int algo, int maximize, double eps, double start_eps, uint32_t max_iterations,
                    uint32_t* person_to_object, uint32_t* object_to_person, double* prices, sla_stats* per_instance_stats,
                    sla_stats* total) {
    if (!ctx) return SLA_ERR_INVALID;
    sla_batch_state* b = ctx->batch;
    if (!b || !b->ready) return fail(ctx, SLA_ERR_STATE, "sla_batch_solve called before a batch was uploaded");
    CU(cudaSetDevice(ctx->device));
    sla::BatchParams bp;
    bp.row_off = b->d_row_off; bp.col_off = b->d_col_off; bp.row_ptr = b->d_row_ptr; bp.cols = b->d_cols; bp.vals = b->d_vals;
    bp.p2o = b->d_p2o; bp.o2p = b->d_o2p; bp.prices = b->d_prices; bp.stats = b->d_stats;
    bp.n_inst = b->n_inst; bp.max_rows = b->max_rows; bp.max_cols = b->max_cols;
    bp.algo = (algo == SLA_ALGO_FORWARD) ? sla::ALGO_FORWARD : sla::ALGO_KHOSLA;
    bp.maximize = maximize ? 1u : 0u;
    bp.max_iterations = max_iterations;
    bp.khosla_scaling = ctx->opt_khosla_scaling ? 1u : 0u;
    bp.eps_in = eps; bp.start_eps_in = start_eps;
    const size_t smem = batch_smem_bytes(b->max_rows, b->max_cols);
    int occ = 1;
    switch (b->lpr) {
        case 1: occ = batch_occupancy_t<1>(smem); break;
        case 2: occ = batch_occupancy_t<2>(smem); break;
        case 4: occ = batch_occupancy_t<4>(smem); break;
        case 8: occ = batch_occupancy_t<8>(smem); break;
        case 16: occ = batch_occupancy_t<16>(smem); break;
        default: occ = batch_occupancy_t<32>(smem); break;
    }
    int grid = ctx->num_sms * occ;
    if ((uint32_t)grid > b->n_inst) grid = (int)b->n_inst;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    int rc;
    switch (b->lpr) {
        case 1: rc = batch_launch_t<1>(ctx, bp, smem, grid); break;
        case 2: rc = batch_launch_t<2>(ctx, bp, smem, grid); break;
        case 4: rc = batch_launch_t<4>(ctx, bp, smem, grid); break;
        case 8: rc = batch_launch_t<8>(ctx, bp, smem, grid); break;
        case 16: rc = batch_launch_t<16>(ctx, bp, smem, grid); break;
        default: rc = batch_launch_t<32>(ctx, bp, smem, grid); break;
    }
    if (rc) return rc;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    b->h_stats.resize(b->n_inst);
    CU(cudaMemcpyAsync(b->h_stats.data(), b->d_stats, (size_t)b->n_inst * sizeof(sla::DevInstStats), cudaMemcpyDeviceToHost,
                       ctx->stream));
    if (person_to_object)
        CU(cudaMemcpyAsync(person_to_object, b->d_p2o, (size_t)b->total_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (object_to_person)
        CU(cudaMemcpyAsync(object_to_person, b->d_o2p, (size_t)b->total_cols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (prices) CU(cudaMemcpyAsync(prices, b->d_prices, (size_t)b->total_cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    sla_stats sum;
    memset(&sum, 0, sizeof sum);
    sum.optimal_soln_found = 1;
    for (uint32_t i = 0; i < b->n_inst; ++i) {
        const sla::DevInstStats& d = b->h_stats[i];
        if (per_instance_stats) {
            sla_stats& o = per_instance_stats[i];
            memset(&o, 0, sizeof o);
            o.num_unassigned = d.num_unassigned; o.nits = d.nits; o.nreductions = d.nreductions;
            o.optimal_soln_found = d.optimal; o.eps = d.eps; o.rounds = d.rounds; o.bids = d.bids; o.bid_arcs = d.bid_arcs;
            o.dropped = d.dropped; o.values_negated = d.values_negated; o.tail_rounds = d.rounds; o.restarts = d.restarts;
        }
        sum.num_unassigned += d.num_unassigned; sum.nits += d.nits; sum.nreductions += d.nreductions;
        sum.optimal_soln_found &= d.optimal; sum.rounds += d.rounds; sum.bids += d.bids; sum.bid_arcs += d.bid_arcs;
        sum.dropped += d.dropped; sum.values_negated += d.values_negated; sum.tail_rounds += d.rounds; sum.restarts += d.restarts;
    }
    sum.kernel_launches = 1;
    cudaEventElapsedTime(&sum.ms_solve, ctx->ev[0], ctx->ev[1]);
    cudaEventElapsedTime(&sum.ms_total, ctx->ev[0], ctx->ev[2]);
    if (total) *total = sum;
    return SLA_OK;
}

}  // extern "C"
