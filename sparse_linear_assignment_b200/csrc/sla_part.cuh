// sla_part.cuh -- one KhoslaSolver instance row-partitioned across ranks (BASELINE.json config 5).
//
// Rank r holds the CSR rows (persons) [row_begin, row_begin + num_rows) and its own bidder queue; the object state
// (prices, owners, packed best-bid words) is replicated.  One synchronous round is
//     sla_part_bid    local bid scan -> local maxima in best[]                      (same kernels as one GPU)
//     all-reduce MAX  over best[] (uint64 words, bit 63 clear => also valid as int64)  [caller, NCCL]
//     sla_part_claim  local winners publish their exact f64 bid in cand[]; losers re-queue
//     all-reduce MAX  over cand[] (f64, -inf = no bid; only one rank writes a finite value per object)
//     sla_part_assign every rank applies ALL winners to its replica (dense sweep), evicted local persons re-queue
// Every rank ends each round with identical replicas, and the result equals the one-GPU solve (and
// oracle/jacobi_model.c) bit for bit, because winners are elected by the same packed words.
// The dense exchange moves 16 bytes per object per round and is communication-bound at cfg5.
#pragma once

namespace sla {

// Local bidders: the winner of object j (its word survived the global MAX) publishes its exact bid; a loser goes
// back to the local queue.
__global__ void __launch_bounds__(kWideThreads) part_claim_kernel(const Params p, double* __restrict__ cand) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (qlen == 0) return;
    const bool identity = h.identity != 0;
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    uint32_t* __restrict__ next_queue = cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[(cur ^ 1u) & 1u];
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (qlen + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        const uint32_t q = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t emit = SLA_DEV_NONE;
        if (q < qlen) {
            const uint32_t j = p.slot_obj[q];
            if (j != SLA_DEV_NONE) {
                const uint32_t i = identity ? q : __ldg(queue + q);
                const double bid = p.slot_bid[q];
                const bool won = (bid == bid) && (__ldcg(p.best + j) == pack_bid(bid, i + h.person_base, h.pbits));
                if (won) cand[j] = bid;
                else emit = i;
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next_len, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit != SLA_DEV_NONE) next_queue[base + __popc(ballot & ((1u << lane) - 1u))] = emit;
        }
    }
}

// Dense sweep over the objects: apply every winner to the replica; local bookkeeping for local persons.
__global__ void __launch_bounds__(kWideThreads) part_apply_kernel(const Params p, double* __restrict__ cand,
                                                                  const uint32_t n_cols, const uint32_t row_begin,
                                                                  const uint32_t n_local) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    uint32_t* __restrict__ next_queue = h.cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[(h.cur ^ 1u) & 1u];
    const unsigned long long pmask = (1ull << h.pbits) - 1ull;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (n_cols + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        const uint32_t j = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t emit = SLA_DEV_NONE;
        if (j < n_cols) {
            const unsigned long long w = p.best[j];
            if (w != 0ull) {
                const uint32_t person = (uint32_t)(pmask - (w & pmask));
                const uint32_t prev = p.o2p[j];
                p.prices[j] = cand[j];
                p.o2p[j] = person;
                p.best[j] = 0ull;
                cand[j] = neg_inf();
                if (person - row_begin < n_local) p.p2o[person - row_begin] = j;
                if (prev != SLA_DEV_NONE && prev - row_begin < n_local) {
                    p.p2o[prev - row_begin] = SLA_DEV_NONE;
                    emit = prev - row_begin;
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next_len, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit != SLA_DEV_NONE) next_queue[base + __popc(ballot & ((1u << lane) - 1u))] = emit;
        }
    }
}

// ---- sparse exchange: ranks trade the lists of their local winners instead of all-reducing M words -------------------
// Entry layout (3 x int64, all-gathered as one tensor): { object, packed word, bits of the exact f64 bid }.

// Local winners (their word is the local maximum of their object) are appended to the send list and their word is
// cleared again; local losers can never win globally and go straight back to the local queue.
__global__ void __launch_bounds__(kWideThreads) part_collect_kernel(const Params p, long long* __restrict__ send,
                                                                    uint32_t* __restrict__ send_count) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (qlen == 0) return;
    const bool identity = h.identity != 0;
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    uint32_t* __restrict__ next_queue = cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[(cur ^ 1u) & 1u];
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t rounds = (qlen + stride - 1) / stride;
    for (uint32_t it = 0; it < rounds; ++it) {
        const uint32_t q = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t emit = SLA_DEV_NONE, obj = SLA_DEV_NONE;
        unsigned long long word = 0ull;
        double bid = 0.0;
        bool win = false;
        if (q < qlen) {
            const uint32_t j = p.slot_obj[q];
            if (j != SLA_DEV_NONE) {
                const uint32_t i = identity ? q : __ldg(queue + q);
                bid = p.slot_bid[q];
                word = pack_bid(bid, i + h.person_base, h.pbits);
                win = (bid == bid) && (__ldcg(p.best + j) == word);
                if (win) { obj = j; p.best[j] = 0ull; }
                else emit = i;
            }
        }
        const uint32_t wb = __ballot_sync(0xffffffffu, win);
        if (wb) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(send_count, (uint32_t)__popc(wb));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (win) {
                long long* e = send + 3ull * (base + __popc(wb & ((1u << lane) - 1u)));
                e[0] = (long long)obj;
                e[1] = (long long)word;
                e[2] = __double_as_longlong(bid);
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next_len, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit != SLA_DEV_NONE) next_queue[base + __popc(ballot & ((1u << lane) - 1u))] = emit;
        }
    }
}

// Pass 1 over every rank's winners: global maximum word per object.
__global__ void __launch_bounds__(kWideThreads) part_sparse_max_kernel(const Params p, const long long* __restrict__ recv,
                                                                       const long long* __restrict__ counts, const uint32_t world,
                                                                       const uint32_t max_count) {
    const unsigned long long total = (unsigned long long)world * max_count;
    for (unsigned long long f = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; f < total;
         f += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(f / max_count), e = (uint32_t)(f % max_count);
        if ((long long)e >= counts[r]) continue;
        const long long* ent = recv + 3ull * f;
        atomicMax(p.best + (uint32_t)ent[0], (unsigned long long)ent[1]);
    }
}

// Pass 2: the global winner of each object is applied to this rank's replica; local persons that lost re-queue.
__global__ void __launch_bounds__(kWideThreads) part_sparse_apply_kernel(const Params p, const long long* __restrict__ recv,
                                                                         const long long* __restrict__ counts, const uint32_t world,
                                                                         const uint32_t max_count, const uint32_t row_begin,
                                                                         const uint32_t n_local) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    uint32_t* __restrict__ next_queue = h.cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[(h.cur ^ 1u) & 1u];
    const unsigned long long pmask = (1ull << h.pbits) - 1ull;
    const int lane = threadIdx.x & 31;
    const unsigned long long total = (unsigned long long)world * max_count;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long rounds = (total + stride - 1) / stride;
    for (unsigned long long it = 0; it < rounds; ++it) {
        const unsigned long long f = it * stride + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
        uint32_t emit = SLA_DEV_NONE;
        if (f < total) {
            const uint32_t r = (uint32_t)(f / max_count), e = (uint32_t)(f % max_count);
            if ((long long)e < counts[r]) {
                const long long* ent = recv + 3ull * f;
                const uint32_t j = (uint32_t)ent[0];
                const unsigned long long w = (unsigned long long)ent[1];
                const uint32_t person = (uint32_t)(pmask - (w & pmask));
                if (__ldcg(p.best + j) == w) {
                    const uint32_t prev = p.o2p[j];
                    p.prices[j] = __longlong_as_double(ent[2]);
                    p.o2p[j] = person;
                    p.best[j] = 0ull;   // losers comparing later see 0 or this word: neither equals theirs
                    if (person - row_begin < n_local) p.p2o[person - row_begin] = j;
                    if (prev != SLA_DEV_NONE && prev - row_begin < n_local) {
                        p.p2o[prev - row_begin] = SLA_DEV_NONE;
                        emit = prev - row_begin;
                    }
                } else if (person - row_begin < n_local) {
                    emit = person - row_begin;
                }
            }
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
        if (ballot) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(next_len, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit != SLA_DEV_NONE) next_queue[base + __popc(ballot & ((1u << lane) - 1u))] = emit;
        }
    }
}

// Round bookkeeping after the apply sweep (one thread).
__global__ void part_control_kernel(const Params p) {
    DevState* st = p.st;
    const uint32_t cur = st->cur & 1u;
    const uint32_t qlen = st->qlen[cur];
    st->rounds += 1;
    st->wide_rounds += 1;
    st->bids += qlen;
    if (st->regular_k) st->bid_arcs += (unsigned long long)qlen * st->regular_k;
    st->qlen[cur] = 0;
    st->cur = cur ^ 1u;
    st->identity = 0;
    st->zero_prices = 0;
}

__global__ void __launch_bounds__(kWideThreads) part_clear_kernel(double* cand, const uint32_t n_cols) {
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_cols; j += gridDim.x * blockDim.x) cand[j] = neg_inf();
}

}  // namespace sla

struct sla_part_state {
    double* d_cand = nullptr;   // exact f64 bid of the winner per object, -inf = no bid (all-reduced with MAX)
    size_t cap_cols = 0;
    // sparse exchange
    long long* d_send = nullptr;      // 3 x int64 per local winner
    long long* d_recv = nullptr;      // world x max_count entries
    long long* d_counts = nullptr;    // winners per rank (all-gathered by the caller), world entries
    uint32_t* d_send_count = nullptr;
    size_t cap_send = 0, cap_recv = 0, cap_world = 0;
    uint32_t row_begin = 0, global_rows = 0;
    bool active = false, first_round = true;
    double eps = 0.0;
    int flip = 0;
    uint32_t launches = 0;
};

extern "C" {

void sla_part_free(sla_ctx* ctx) {
    if (!ctx || !ctx->part) return;
    cudaFree(ctx->part->d_cand);
    cudaFree(ctx->part->d_send); cudaFree(ctx->part->d_recv); cudaFree(ctx->part->d_counts); cudaFree(ctx->part->d_send_count);
    delete ctx->part;
    ctx->part = nullptr;
}

int sla_part_local_value_range(sla_ctx* ctx, double* w_min, double* w_max, double* first_value) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->has_csr) return fail(ctx, SLA_ERR_STATE, "no CSR shard uploaded");
    if (w_min) *w_min = ctx->v_min;
    if (w_max) *w_max = ctx->v_max;
    if (first_value) *first_value = ctx->first_value;
    return SLA_OK;
}

int sla_part_begin(sla_ctx* ctx, int algo, int maximize, uint32_t row_begin, uint32_t global_rows, double eps,
                   double global_w_min, double global_w_max, double global_first_value) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->has_csr) return fail(ctx, SLA_ERR_STATE, "sla_part_begin called before the CSR shard was uploaded");
    if (algo != SLA_ALGO_KHOSLA) return fail(ctx, SLA_ERR_INVALID, "the row-partitioned engine implements KhoslaSolver only");
    if ((uint64_t)row_begin + ctx->n_rows > global_rows) return fail(ctx, SLA_ERR_INVALID, "shard exceeds global_rows");
    CU(cudaSetDevice(ctx->device));
    if (!ctx->part) ctx->part = new sla_part_state();
    sla_part_state* ps = ctx->part;
    if (ctx->n_cols > ps->cap_cols) {
        int rc = dev_alloc(ctx, &ps->d_cand, ctx->n_cols);
        if (rc) return rc;
        ps->cap_cols = ctx->n_cols;
    }
    const uint32_t N = ctx->n_rows, M = ctx->n_cols;
    // sign normalisation of the GLOBAL instance (solver.rs:207-216): decided by the very first value (rank 0's)
    const bool flip = (maximize != 0) != (global_first_value >= 0.0);
    ctx->dev_sign = flip ? -1 : 1;
    const double w_min = flip ? -global_w_max : global_w_min, w_max = flip ? -global_w_min : global_w_max;
    DevState s;
    memset(&s, 0, sizeof s);
    s.qlen[0] = N;
    s.identity = 1;
    s.zero_prices = 1;
    s.algo = ALGO_KHOSLA;
    s.pbits = person_bits(global_rows);
    s.tail_max = 0;                       // wide kernels only: every round needs the exchange
    s.skip_zero = (uint32_t)ctx->opt_skip_zero;
    s.sign_flip = flip ? 0x80000000u : 0u;
    s.n_rows = N;
    s.n_cols = M;
    s.person_base = row_begin;
    s.regular_k = use_regular(ctx) ? ctx->regular_k : 0u;
    s.safety_rounds_left = 1ull << 40;
    s.max_iterations = 0xFFFFFFFFu;
    const double m = (double)M;
    s.eps = std::isnan(eps) ? 1.0 / m : eps;                        // ksparse.rs:162-169
    s.threshold = (m / 2.0) * (w_max - w_min + s.eps);               // ksparse.rs:181
    *ctx->h_state = s;
    CU(cudaMemcpyAsync(ctx->d_state, ctx->h_state, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    const Params p = make_params(ctx);
    init_solve_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, N, M, 1);
    part_clear_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(ps->d_cand, M);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ps->row_begin = row_begin;
    ps->global_rows = global_rows;
    ps->active = true;
    ps->first_round = true;
    ps->eps = s.eps;
    ps->flip = flip ? 1 : 0;
    ps->launches = 2;
    ctx->has_solution = false;
    ctx->best_dirty = true;
    return SLA_OK;
}

int sla_part_buffers(sla_ctx* ctx, void** d_best_words, void** d_price_candidates, uint64_t* num_words) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    if (d_best_words) *d_best_words = ctx->d_best;
    if (d_price_candidates) *d_price_candidates = ctx->part->d_cand;
    if (num_words) *num_words = ctx->n_cols;
    return SLA_OK;
}

int sla_part_bid(sla_ctx* ctx) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    launch_one(ctx, p, 0, ctx->part->first_round && ctx->opt_skip_zero != 0);
    ctx->part->first_round = false;
    ctx->part->launches += 1;
    CU(cudaGetLastError());
    return SLA_OK;
}

int sla_part_claim(sla_ctx* ctx) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    part_claim_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->part->d_cand);
    ctx->part->launches += 1;
    CU(cudaGetLastError());
    return SLA_OK;
}

int sla_part_assign(sla_ctx* ctx, uint32_t* local_queue_len, uint32_t* local_dropped) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    part_apply_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->part->d_cand, ctx->n_cols, ctx->part->row_begin,
                                                                       ctx->n_rows);
    part_control_kernel<<<1, 1, 0, ctx->stream>>>(p);
    ctx->part->launches += 2;
    int rc = poll_state(ctx);
    if (rc) return rc;
    CU(cudaGetLastError());
    const DevState& f = *ctx->h_state;
    if (local_queue_len) *local_queue_len = f.qlen[f.cur & 1u];
    if (local_dropped) *local_dropped = f.dropped;
    return SLA_OK;
}

int sla_part_finish(sla_ctx* ctx, uint32_t* person_to_object, uint32_t* object_to_person, double* prices, sla_stats* stats) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    CU(cudaSetDevice(ctx->device));
    int rc = poll_state(ctx);
    if (rc) return rc;
    const DevState f = *ctx->h_state;
    if (person_to_object)
        CU(cudaMemcpyAsync(person_to_object, ctx->d_p2o, (size_t)ctx->n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (object_to_person)
        CU(cudaMemcpyAsync(object_to_person, ctx->d_o2p, (size_t)ctx->n_cols * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (prices) CU(cudaMemcpyAsync(prices, ctx->d_prices, (size_t)ctx->n_cols * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->num_unassigned = f.dropped;            // local share; the caller sums over ranks
        stats->nits = (uint32_t)f.bids;
        stats->eps = ctx->part->eps;
        stats->rounds = f.rounds;
        stats->bids = f.bids;
        stats->bid_arcs = f.bid_arcs;
        stats->dropped = f.dropped;
        stats->values_negated = (uint32_t)ctx->part->flip;
        stats->wide_rounds = f.wide_rounds;
        stats->kernel_launches = ctx->part->launches;
    }
    ctx->part->active = false;
    ctx->has_solution = true;
    ctx->best_dirty = false;
    return SLA_OK;
}

// ---- sparse exchange entry points ----
int sla_part_sparse_buffers(sla_ctx* ctx, int world, void** d_send, void** d_recv, void** d_counts, uint64_t* send_capacity) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active) return fail(ctx, SLA_ERR_STATE, "sla_part_begin has not been called");
    if (world < 1) return fail(ctx, SLA_ERR_INVALID, "world must be >= 1");
    CU(cudaSetDevice(ctx->device));
    sla_part_state* ps = ctx->part;
    int rc;
    // every rank may send up to the largest shard (shards differ by at most one row): size by that bound
    const size_t cap = (size_t)(ps->global_rows + (uint32_t)world - 1) / (size_t)world + 1;
    if (cap > ps->cap_send) {
        if ((rc = dev_alloc(ctx, &ps->d_send, 3 * cap))) return rc;
        ps->cap_send = cap;
    }
    if ((size_t)world * cap > ps->cap_recv) {
        if ((rc = dev_alloc(ctx, &ps->d_recv, 3 * (size_t)world * cap))) return rc;
        ps->cap_recv = (size_t)world * cap;
    }
    if ((size_t)world > ps->cap_world) {
        if ((rc = dev_alloc(ctx, &ps->d_counts, (size_t)world))) return rc;
        ps->cap_world = (size_t)world;
    }
    if (!ps->d_send_count && (rc = dev_alloc(ctx, &ps->d_send_count, 4))) return rc;
    if (d_send) *d_send = ps->d_send;
    if (d_recv) *d_recv = ps->d_recv;
    if (d_counts) *d_counts = ps->d_counts;
    if (send_capacity) *send_capacity = ps->cap_send;
    return SLA_OK;
}

int sla_part_collect(sla_ctx* ctx, uint32_t* local_winners) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active || !ctx->part->d_send) return fail(ctx, SLA_ERR_STATE, "sla_part_sparse_buffers has not been called");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    sla_part_state* ps = ctx->part;
    CU(cudaMemsetAsync(ps->d_send_count, 0, sizeof(uint32_t), ctx->stream));
    part_collect_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ps->d_send, ps->d_send_count);
    CU(cudaMemcpyAsync(ctx->h_scratch, ps->d_send_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ps->launches += 1;
    if (local_winners) *local_winners = ctx->h_scratch[0];
    return SLA_OK;
}

int sla_part_apply_sparse(sla_ctx* ctx, int world, uint32_t max_count, uint32_t* local_queue_len, uint32_t* local_dropped) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!ctx->part || !ctx->part->active || !ctx->part->d_recv) return fail(ctx, SLA_ERR_STATE, "sla_part_sparse_buffers has not been called");
    if ((size_t)world * max_count > ctx->part->cap_recv) return fail(ctx, SLA_ERR_INVALID, "gathered list exceeds the receive buffer");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    sla_part_state* ps = ctx->part;
    if (max_count) {
        part_sparse_max_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ps->d_recv, ps->d_counts, (uint32_t)world, max_count);
        part_sparse_apply_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ps->d_recv, ps->d_counts, (uint32_t)world, max_count,
                                                                                  ps->row_begin, ctx->n_rows);
        ps->launches += 2;
    }
    part_control_kernel<<<1, 1, 0, ctx->stream>>>(p);
    ps->launches += 1;
    int rc = poll_state(ctx);
    if (rc) return rc;
    CU(cudaGetLastError());
    const DevState& f = *ctx->h_state;
    if (local_queue_len) *local_queue_len = f.qlen[f.cur & 1u];
    if (local_dropped) *local_dropped = f.dropped;
    return SLA_OK;
}

}  // extern "C"
