// sla_mesh.cuh -- one KhoslaSolver instance over G GPUs of one NVLink / NVSwitch domain (BASELINE.json config 5).
//
// Persons (CSR rows) are partitioned by row over the ranks, objects by contiguous ranges of S = 2^shift ids over the
// SAME ranks: rank g owns {price, owner} (one 16-byte cell) and the packed bid word of objects [g S, (g+1) S) in its own
// HBM.  Every rank maps every other rank's "mesh block" (peer mapping: cudaIpc* between processes, plain pointers inside
// one process), and all exchange happens INSIDE the round kernels, as loads and stores on those peer pointers:
//
//   K1 bid      every queued local person scans its CSR row (the single-GPU choice rule, ksparse.rs:199-227); prices of
//               objects another rank owns are gathered straight out of that rank's HBM; the bid {object, person, exact
//               f64 bid} is staged in shared memory, sorted by owner, and PUSHED with coalesced stores into the owner's
//               inbox over NVLink.  Last block: entry counts to the owners, flag barrier B1.
//   K2 max      owner side, local: packed 64-bit word of every received bid -> atomicMax on the object's word.  Also the
//               termination test: the sum of the ranks' queue lengths of this round == 0 ends the solve everywhere.
//   K3 resolve  owner side: the entry whose word survived wins -- price := its exact bid, owner := its person, the
//               previous owner is pushed into the evict inbox of the rank that holds that person; one reply bit per entry
//               is stored back into the bidder's rank (32 entries = one word).  Last block: counts, barrier B2.
//   K4 finish   bidder side: winners record their object, losers and the evicted persons received from the owners form
//               the next queue.  Last block: round accounting, this rank's next queue length to every rank (it travels
//               under the next B1).
//
// A barrier is one 32-bit epoch per (receiver, sender) pair: every block fences its stores at system scope before it
// takes its ticket; the last block of the producing kernel stores the epoch into every rank's flag word and then WAITS
// for the other ranks' epochs itself, so the kernel boundary behind it is the barrier and nobody else polls.  Rounds are
// captured into CUDA graphs; the host only polls `done` -- no collective, no host synchronisation per round.
//
// A Jacobi round does not depend on the order of its bids, winners are elected by the same packed words, and prices
// are the winners' exact f64 bids: the result equals the one-GPU solve (and oracle/jacobi_model.c) bit for bit.
#pragma once

namespace sla {

constexpr uint32_t kMeshChunkRows = 1024;    // bidders one block stages in shared memory before it pushes them
constexpr uint32_t kMeshPruneMinQueue = 4096; // local bidders from which the gathering scan prunes by value bound

struct alignas(16) BidEntry {
    uint32_t obj_local;      // object id relative to the owner's first object
    uint32_t person;         // global person id
    double bid;              // exact f64 bid
};
static_assert(sizeof(BidEntry) == 16, "one bid entry is one 128-bit store");

// The part of a rank's mesh block that peers write: flags and counts.  One 128-byte line per array.
struct alignas(128) MeshMailbox {
    uint32_t flags[32];              // flags[src]: last barrier epoch rank `src` has signalled to this rank
    uint32_t bid_count[32];          // bid entries rank `src` pushed into this rank's inbox in the current round
    uint32_t evict_count[32];        // evicted persons rank `src` pushed in the current round
    unsigned long long next_total[16];   // next_total[src]: length of rank src's next queue (their sum == 0: solved)
};

// Everything a mesh kernel needs beyond Params; passed by value as a __grid_constant__ kernel parameter.  `view` must
// stay the first member: the scan helpers receive its address in place of the price array (ld_price<PRICE_MESH>).
struct MeshParams {
    MeshView view;                               // cells of every rank, shift, mask
    MeshMailbox* box[kMeshMaxRanks];             // mailbox of every rank (own included)
    BidEntry* bid_out[kMeshMaxRanks];            // bid_out[g]: the region of rank g's bid inbox reserved for this rank
    uint32_t* reply_out[kMeshMaxRanks];          // reply_out[a]: the region of rank a's reply words written by this rank
    uint32_t* evict_out[kMeshMaxRanks];          // evict_out[b]: the region of rank b's evict inbox reserved for this rank
    ObjCell* my_cells;
    unsigned long long* my_best;                 // [2^shift] packed bid word of the current round per owned object (0 = none)
    BidEntry* my_bid_in;                         // [world][cap_bid]   entries received, by sender
    uint32_t* my_reply_in;                       // [world][cap_words] reply words received, by owner
    uint32_t* my_evict_in;                       // [world][cap_evict] evicted persons received, by owner
    uint32_t* slot_pos;                          // per queue slot: position of its entry in the owner's inbox region
    uint32_t* out_cnt;                           // [world] entries pushed to each owner this round (local counter)
    uint32_t* ev_cnt;                            // [world] evictions pushed to each rank this round (local counter)
    uint32_t* tickets;                           // [4] last-block tickets of the four kernels
    uint32_t rank, world;
    uint32_t cap_bid, cap_words, cap_evict;
    uint32_t my_objects;                         // objects this rank owns
    uint32_t row_begin[kMeshMaxRanks + 1];       // global person id of every rank's first row (+ total)
    unsigned long long timeout_ns;               // a barrier that waits longer gives up (DevState::mesh_error)
    // Where a barrier is waited for.  0 (concurrent ranks, one per GPU): in the tail of the PRODUCING kernel -- its last
    // block signals and then waits for the peers' signals, so the kernel boundary behind it is the barrier and nobody
    // else polls (a thousand blocks polling one L2 line delay the very store they wait for).  1 (lockstep: ranks
    // stepped one kernel at a time, possibly on one GPU): at the start of the CONSUMING kernel, where the flags are
    // already there -- a producer that waited in its tail would wait for kernels that have not been launched yet.
    uint32_t wait_at_start;
    uint32_t tail_engine;                        // option: hand the short rounds to mesh_tail_kernel
    unsigned long long* timeline;                // [kMeshTimelineRounds][kMeshTimelineSlots] globaltimer stamps (development aid), or nullptr
};
constexpr uint32_t kMeshTimelineRounds = 64, kMeshTimelineSlots = 12;

// ---- barrier ----------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Every block: wait until all ranks have signalled `epoch` to this rank.  Returns false when the wait gave up (a peer
// died or never launched): the caller leaves the kernel; DevState::mesh_error / done stop everything behind it.
__device__ __forceinline__ bool mesh_wait(const MeshParams& mp, DevState* st, const uint32_t epoch) {
    __shared__ uint32_t s_fail;
    if (threadIdx.x == 0) s_fail = 0u;
    __syncthreads();
    if (threadIdx.x < mp.world) {
        const volatile uint32_t* f = &mp.box[mp.rank]->flags[threadIdx.x];
        const unsigned long long t0 = global_timer_ns();
        uint32_t spins = 0;
        while ((int)(*f - epoch) < 0) {
            if ((++spins & 1023u) == 0u) {
                if (global_timer_ns() - t0 > mp.timeout_ns || ((volatile DevState*)st)->mesh_error) { s_fail = 1u; break; }
            }
            __nanosleep(64);
        }
        __threadfence_system();      // acquire: what the signalling ranks stored before their flags is visible from here on
    }
    __syncthreads();
    if (s_fail) {
        if (threadIdx.x == 0) { ((volatile DevState*)st)->mesh_error = 1u; ((volatile DevState*)st)->done = 1u; }
        return false;
    }
    return true;
}

// One thread, after the data this rank produced has been fenced: tell every rank (this one included).
__device__ __forceinline__ void mesh_signal(const MeshParams& mp, const uint32_t epoch) {
    __threadfence_system();
    for (uint32_t g = 0; g < mp.world; ++g) st_release_sys_u32(&mp.box[g]->flags[mp.rank], epoch);
}

// Last block of a producing kernel (all its threads): signal `epoch` to every rank and -- unless the consumers wait for
// themselves (wait_at_start) -- wait until every rank has signalled it to this one.  Thread g talks to rank g.
__device__ __forceinline__ void mesh_barrier_tail(const MeshParams& mp, DevState* st, const uint32_t epoch) {
    __threadfence_system();
    if (threadIdx.x < mp.world) {
        *reinterpret_cast<volatile uint32_t*>(&mp.box[threadIdx.x]->flags[mp.rank]) = epoch;
        if (!mp.wait_at_start) {
            const volatile uint32_t* f = &mp.box[mp.rank]->flags[threadIdx.x];
            const unsigned long long t0 = global_timer_ns();
            uint32_t spins = 0;
            while ((int)(*f - epoch) < 0) {
                if ((++spins & 255u) == 0u && global_timer_ns() - t0 > mp.timeout_ns) {
                    ((volatile DevState*)st)->mesh_error = 1u;
                    ((volatile DevState*)st)->done = 1u;
                    break;
                }
            }
            __threadfence_system();  // acquire
        }
    }
}

__device__ __forceinline__ void mesh_stamp(const MeshParams& mp, DevState* st, const uint32_t slot) {
    if (mp.timeline) {
        const uint32_t r = ((volatile DevState*)st)->mesh_round;
        if (r < kMeshTimelineRounds) mp.timeline[r * kMeshTimelineSlots + slot] = global_timer_ns();
    }
}

// Last-block detection (classic threadfence reduction): true in exactly one thread of the grid, after every block's
// stores are visible to it.
// `remote`: this block has stored into another rank's memory since its last fence (block-uniform or not: any thread's
// word counts).  Only such blocks pay for a system-scope fence (it waits for the acknowledgements from across NVLink);
// the others order their local stores at GPU scope.
__device__ __forceinline__ bool mesh_last_block(uint32_t* ticket, const bool remote, const uint32_t participants) {
    __shared__ uint32_t s_is_last;
    const int any_remote = __syncthreads_or(remote ? 1 : 0);
    if (threadIdx.x == 0) {
        if (any_remote) __threadfence_system(); else __threadfence();
        const uint32_t t = atomicAdd(ticket, 1u);
        s_is_last = (t == participants - 1u) ? 1u : 0u;
        if (s_is_last) {
            __threadfence();
            *ticket = 0u;
        }
    }
    __syncthreads();
    return s_is_last != 0u;      // block-uniform: every thread of the last block gets true
}

__device__ __forceinline__ uint32_t mesh_person_rank(const MeshParams& mp, const uint32_t person) {
    uint32_t g = 0;
    while (g + 1u < mp.world && person >= mp.row_begin[g + 1u]) ++g;
    return g;
}

// Loads past the L1 for words other ranks store into this rank's HBM.
__device__ __forceinline__ uint32_t ld_cv_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// =============================================================================================================
// K1: bid scan + push of the bids into the owners' inboxes (fused compute + communication).
// =============================================================================================================
struct MeshStage {
    BidEntry ent[kMeshChunkRows];        // arrival order
    BidEntry sorted[kMeshChunkRows];     // grouped by owner rank
    uint32_t slot[kMeshChunkRows];       // queue slot of the entry (arrival order)
    uint32_t where[kMeshChunkRows];      // owner << 24 | rank within the owner's entries of this chunk (arrival order)
    uint32_t slot_sorted[kMeshChunkRows];
    uint32_t cnt[kMeshMaxRanks], off[kMeshMaxRanks + 1], base[kMeshMaxRanks];
    uint32_t n, dropped;
    unsigned long long arcs;
};

__device__ __forceinline__ void mesh_stage_reset(MeshStage& s) {
    if (threadIdx.x < kMeshMaxRanks) s.cnt[threadIdx.x] = 0u;
    if (threadIdx.x == 0) s.n = 0u;
}

// One finished bid of queue slot q (lane 0 of its group).
__device__ __forceinline__ void mesh_stage_bid(MeshStage& s, const Params& p, const MeshParams& mp, const Bid& r,
                                               const uint32_t q, const uint32_t person_global) {
    if (r.dropped) {
        p.slot_obj[q] = SLA_DEV_NONE;
        atomicAdd(&s.dropped, 1u);
        return;
    }
    if (!(r.bid == r.bid)) {            // NaN never bids (symmetric.rs:394): the person stays in the queue
        p.slot_obj[q] = r.obj;
        mp.slot_pos[q] = SLA_DEV_NONE;
        return;
    }
    p.slot_obj[q] = r.obj;
    const uint32_t g = r.obj >> mp.view.shift;
    const uint32_t e = atomicAdd(&s.n, 1u);
    const uint32_t rk = atomicAdd(&s.cnt[g], 1u);
    BidEntry en;
    en.obj_local = r.obj & mp.view.mask;
    en.person = person_global;
    en.bid = r.bid;
    s.ent[e] = en;
    s.slot[e] = q;
    s.where[e] = (g << 24) | rk;
}

// After a chunk: reserve room in every owner's inbox region -- ONE global atomic per owner and chunk: atomics on one
// address serialise at ~8 ns apiece -- group the staged entries by owner in shared memory, and store them out:
// consecutive threads write consecutive 16-byte entries of one owner's region, so what crosses NVLink are whole
// 128-byte lines whatever the number of ranks (in arrival order a warp's 32 entries would split into 64-byte pieces for
// eight owners: the first round at 8 GPUs was bound by those packets).
__device__ __forceinline__ void mesh_stage_flush(MeshStage& s, const MeshParams& mp) {
    __syncthreads();
    const uint32_t n = s.n;
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (uint32_t g = 0; g < mp.world; ++g) { s.off[g] = acc; acc += s.cnt[g]; }
        for (uint32_t g = mp.world; g <= kMeshMaxRanks; ++g) s.off[g] = acc;
    }
    if (threadIdx.x < mp.world) s.base[threadIdx.x] = s.cnt[threadIdx.x] ? atomicAdd(&mp.out_cnt[threadIdx.x], s.cnt[threadIdx.x]) : 0u;
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
        const uint32_t w = s.where[e], g = w >> 24, at = s.off[g] + (w & 0xFFFFFFu);
        s.sorted[at] = s.ent[e];
        s.slot_sorted[at] = s.slot[e];
    }
    __syncthreads();
    for (uint32_t t = threadIdx.x; t < n; t += blockDim.x) {
        uint32_t g = 0;
#pragma unroll
        for (int k = 1; k < kMeshMaxRanks; ++k) g += (t >= s.off[k]) ? 1u : 0u;
        const uint32_t pos = s.base[g] + (t - s.off[g]);
        mp.bid_out[g][pos] = s.sorted[t];
        mp.slot_pos[s.slot_sorted[t]] = pos;
    }
    __syncthreads();
    mesh_stage_reset(s);
    __syncthreads();
}

// Blocks of a grid that take part in the last-block ticket of a kernel: those that had work (at least one; block 0
// stands in when nobody had).  The others have stored nothing and leave at once.
__device__ __forceinline__ uint32_t mesh_working_blocks(const uint32_t work_units) {
    const uint32_t w = work_units < gridDim.x ? work_units : gridDim.x;
    return w ? w : 1u;
}

// Start of K1.  Nothing to wait for: the queue is this rank's own (K4 of the previous round), and the prices it gathers
// were written by the owners before the barrier B2 this rank has already passed.
__device__ __forceinline__ bool mesh_round_begin(const Params& p, const MeshParams& mp, const HotState& h) {
    if (h.done) return false;
    if (((volatile DevState*)p.st)->mesh_error) return false;
    return true;
}

// End of K1 (last block): counts to the owners, barrier B1.
// Rows a block stages per pass over its share of the queue: the whole staging area for the long rounds (one global atomic
// per owner and 1,024 bids), less for the short ones so that the queue still spreads over the whole grid.
__device__ __forceinline__ uint32_t mesh_chunk_rows(const uint32_t qlen) {
    uint32_t c = (qlen + gridDim.x - 1u) / gridDim.x;
    c = (c + 255u) & ~255u;
    return c < 256u ? 256u : (c > kMeshChunkRows ? kMeshChunkRows : c);
}

__device__ __forceinline__ void mesh_bid_end(const Params& p, const MeshParams& mp, MeshStage& s, const uint32_t qlen) {
    DevState* st = p.st;
    const uint32_t chunk_rows = mesh_chunk_rows(qlen);
    const uint32_t working = mesh_working_blocks((qlen + chunk_rows - 1u) / chunk_rows);
    if (blockIdx.x >= working) return;
    if (threadIdx.x == 0) {
        if (s.dropped) atomicAdd(&st->dropped, s.dropped);
        if (s.arcs) atomicAdd(&st->bid_arcs, s.arcs);
    }
    if (mesh_last_block(&mp.tickets[0], qlen != 0u, working)) {
        const uint32_t epoch = ((volatile DevState*)st)->mesh_epoch;
        if (threadIdx.x < mp.world) {
            const uint32_t g = threadIdx.x;
            const uint32_t c = *reinterpret_cast<volatile uint32_t*>(&mp.out_cnt[g]);
            *reinterpret_cast<volatile uint32_t*>(&mp.box[g]->bid_count[mp.rank]) = c;
            *reinterpret_cast<volatile uint32_t*>(&mp.out_cnt[g]) = 0u;
        }
        if (threadIdx.x == 0) mesh_stamp(mp, st, 1);
        mesh_barrier_tail(mp, st, epoch + 1u);
        if (threadIdx.x == 0) mesh_stamp(mp, st, 2);
    }
}

// Uniform-degree CSR (K % 8 == 0).  ZERO: first round of a solve, all prices exactly 0 (no gather); NARROW: values
// read from their u16 mirror.  Two rows per lane group in flight.
template <int LPR8, bool ZERO, bool NARROW, bool PERSIST>
__device__ __forceinline__ void mesh_bid_body(const Params& p, const MeshParams& mp, const HotState& h, MeshStage& s) {
    if (!mesh_round_begin(p, mp, h)) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) mesh_stamp(mp, p.st, 0);
    const uint32_t cur = h.cur, qlen = h.qlen[cur & 1u];
    const bool identity = h.identity != 0;
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    const uint32_t K = h.regular_k, sign_flip = h.sign_flip, algo = h.algo, base_person = h.person_base;
    const double eps = h.eps, thr = h.threshold;
    const double* mesh_prices = reinterpret_cast<const double*>(&mp);     // ld_price<PRICE_MESH> reads mp.view
    constexpr int MODE = ZERO ? PRICE_ZERO : (PERSIST ? PRICE_MESH_CG : PRICE_MESH);
    constexpr int GPB = kWideThreads / LPR8;
    constexpr int U = 2;
    const int lane = threadIdx.x % LPR8;
    const uint32_t group = threadIdx.x / LPR8;
    const uint32_t keyflip = sign_flip ? 0xFFFFu : 0u;

    // bound-pruned gather (scan_row_pruned): here the gathers it saves cross NVLink -- from 4 Ki bidders on this rank
    const bool prune = !ZERO && K <= 8u * LPR8 && qlen >= kMeshPruneMinQueue && p.st->prune_ok != 0u;

    mesh_stage_reset(s);
    if (threadIdx.x == 0) { s.dropped = 0u; s.arcs = 0ull; }
    __syncthreads();
    const uint32_t chunk_rows = mesh_chunk_rows(qlen);
    for (uint32_t chunk = blockIdx.x * chunk_rows; chunk < qlen; chunk += gridDim.x * chunk_rows) {
        const uint32_t chunk_end = (chunk + chunk_rows < qlen) ? chunk + chunk_rows : qlen;
        for (uint32_t r0 = chunk; r0 < chunk_end; r0 += GPB * U) {
            uint32_t q[U], i[U];
            bool valid[U];
            Choice c[U];
            KeyChoice kc[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                q[u] = r0 + (uint32_t)u * GPB + group;
                valid[u] = q[u] < chunk_end;
                i[u] = 0;
                choice_init(c[u]);
                key_choice_init(kc[u]);
                if (valid[u]) {
                    i[u] = identity ? q[u] : (PERSIST ? __ldcg(queue + q[u]) : __ldg(queue + q[u]));
                    const uint32_t a = i[u] * K;
                    if (!prune)
                    for (uint32_t off = 8u * (uint32_t)lane; off < K; off += 8u * LPR8) {
                        if (ZERO && NARROW) scan8_keys(kc[u], p.cols, p.vals16, a + off, off, keyflip);
                        else if (NARROW) scan8_narrow<MODE>(c[u], p.cols, p.vals16, mesh_prices, a + off, sign_flip);
                        else scan8<MODE>(c[u], p.cols, p.vals, mesh_prices, a + off, sign_flip);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (ZERO && NARROW) {
                    key_choice_group_reduce<LPR8>(kc[u]);
                    if (valid[u] && lane == 0) key_choice_to_f64(c[u], kc[u], i[u] * K, keyflip, sign_flip);
                } else if (!ZERO && prune) {
                    scan_row_pruned<LPR8, MODE, NARROW>(c[u], p, mesh_prices, i[u], valid[u], K, sign_flip, lane);
                } else {
                    choice_group_reduce<LPR8>(c[u]);
                }
                if (valid[u] && lane == 0) {
                    const Bid r = make_bid<MODE>(c[u], algo, eps, thr, mesh_prices);
                    mesh_stage_bid(s, p, mp, r, q[u], i[u] + base_person);
                }
            }
        }
        mesh_stage_flush(s, mp);
    }
    mesh_bid_end(p, mp, s, qlen);
}

template <int LPR8, bool ZERO, bool NARROW>
__global__ void __launch_bounds__(kWideThreads, 3) mesh_bid_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    __shared__ MeshStage s;
    const HotState h = load_hot(p.st);
    mesh_bid_body<LPR8, ZERO, NARROW, false>(p, mp, h, s);
}

// Ragged CSR: aligned 4-arc chunks, masked ends, extents from row_ptr (the layout of bid_wide_kernel).
template <int LPR, bool PERSIST>
__device__ __forceinline__ void mesh_bid_ragged_body(const Params& p, const MeshParams& mp, const HotState& h, MeshStage& s) {
    if (!mesh_round_begin(p, mp, h)) return;
    const uint32_t cur = h.cur, qlen = h.qlen[cur & 1u];
    const bool identity = h.identity != 0;
    const bool zero = (h.zero_prices != 0) && (h.skip_zero != 0);
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    const uint32_t sign_flip = h.sign_flip, algo = h.algo, base_person = h.person_base;
    const double eps = h.eps, thr = h.threshold;
    const double* mesh_prices = reinterpret_cast<const double*>(&mp);
    constexpr int GPB = kWideThreads / LPR;
    const int lane = threadIdx.x % LPR;
    const uint32_t group = threadIdx.x / LPR;

    mesh_stage_reset(s);
    if (threadIdx.x == 0) { s.dropped = 0u; s.arcs = 0ull; }
    __syncthreads();
    const uint32_t chunk_rows = mesh_chunk_rows(qlen);
    for (uint32_t chunk = blockIdx.x * chunk_rows; chunk < qlen; chunk += gridDim.x * chunk_rows) {
        const uint32_t chunk_end = (chunk + chunk_rows < qlen) ? chunk + chunk_rows : qlen;
        for (uint32_t r0 = chunk; r0 < chunk_end; r0 += GPB) {
            const uint32_t q = r0 + group;
            const bool valid = q < chunk_end;
            uint32_t i = 0, a = 0, b = 0;
            if (valid) {
                i = identity ? q : (PERSIST ? __ldcg(queue + q) : __ldg(queue + q));
                a = __ldg(p.row_ptr + i);
                b = __ldg(p.row_ptr + i + 1);
            }
            constexpr int GMODE = PERSIST ? PRICE_MESH_CG : PRICE_MESH;
            Choice c;
            choice_init(c);
            if (zero) scan_row<LPR, PRICE_ZERO>(c, p.cols, p.vals, mesh_prices, a, b, sign_flip, lane);
            else scan_row<LPR, GMODE>(c, p.cols, p.vals, mesh_prices, a, b, sign_flip, lane);
            choice_group_reduce<LPR>(c);
            if (valid && lane == 0) {
                const Bid r = zero ? make_bid<PRICE_ZERO>(c, algo, eps, thr, mesh_prices)
                                   : make_bid<GMODE>(c, algo, eps, thr, mesh_prices);
                atomicAdd(&s.arcs, (unsigned long long)(b - a));
                mesh_stage_bid(s, p, mp, r, q, i + base_person);
            }
        }
        mesh_stage_flush(s, mp);
    }
    mesh_bid_end(p, mp, s, qlen);
}

template <int LPR>
__global__ void __launch_bounds__(kWideThreads, 3) mesh_bid_ragged_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    __shared__ MeshStage s;
    const HotState h = load_hot(p.st);
    mesh_bid_ragged_body<LPR, false>(p, mp, h, s);
}

// =============================================================================================================
// K2: owner side, pass 1 -- the maximum packed word per object over everything that arrived (local atomics only).
// =============================================================================================================
constexpr uint32_t kMeshTailPerRank = 512;   // average bidders per rank from which the persistent tail engine takes over (two
                                             // passes of its one block; 2.8 k per rank cost it 266 us a round at 2 GPUs)

__device__ __forceinline__ void mesh_max_body(const Params& p, const MeshParams& mp, const HotState& h) {
    DevState* st = p.st;
    if (h.done) return;
    const uint32_t epoch = ((volatile DevState*)st)->mesh_epoch, round = ((volatile DevState*)st)->mesh_round;
    if (mp.wait_at_start && !mesh_wait(mp, st, epoch + 1u)) return;
    if (((volatile DevState*)st)->mesh_error) return;
    if (round > 1u) {
        // the ranks' queue lengths of THIS round (published by their K4 of the previous one, ordered by B1): all zero means
        // nobody bid -- the solve ends here on every rank, in the same round.  `done` is set by one thread, after every block
        // has read the state (ticket), so that no block of this very launch can take the `h.done` exit while others go on.
        unsigned long long total = 0, part[kMeshMaxRanks];
#pragma unroll
        for (int g = 0; g < kMeshMaxRanks; ++g)           // all loads in flight together
            part[g] = ((uint32_t)g < mp.world) ? __ldcg(&mp.box[mp.rank]->next_total[g]) : 0ull;
#pragma unroll
        for (int g = 0; g < kMeshMaxRanks; ++g) total += part[g];
        if (total == 0ull) {
            if (mesh_last_block(&mp.tickets[1], false, gridDim.x) && threadIdx.x == 0) {
                // this round's B1 has been signalled by every rank: the next solve must start behind it
                ((volatile DevState*)st)->mesh_epoch = epoch + 2u;
                ((volatile DevState*)st)->done = 1u;
            }
            return;
        }
        // short rounds from here on: the one-block persistent engine behind this round's finish kernel runs them without
        // kernel boundaries (every rank takes the same decision from the same numbers)
        if (total <= (unsigned long long)kMeshTailPerRank * mp.world && mp.tail_engine && !mp.wait_at_start && blockIdx.x == 0 && threadIdx.x == 0 &&
            ((volatile DevState*)st)->mesh_tail == 0u) {
            ((volatile DevState*)st)->mesh_switch_round = round;
            ((volatile DevState*)st)->mesh_tail = 1u;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) mesh_stamp(mp, st, 3);
    const uint32_t pbits = h.pbits;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    uint32_t cnts[kMeshMaxRanks];
#pragma unroll
    for (int a = 0; a < kMeshMaxRanks; ++a) cnts[a] = ((uint32_t)a < mp.world) ? __ldcg(&mp.box[mp.rank]->bid_count[a]) : 0u;
#pragma unroll
    for (int a = 0; a < kMeshMaxRanks; ++a) {
        const uint32_t cnt = cnts[a];
        const BidEntry* in = mp.my_bid_in + (size_t)a * mp.cap_bid;
        for (uint32_t e = tid; e < cnt; e += stride) {
            const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(in + e));      // stored by another GPU: past the L1
            const double bid = __hiloint2double((int)raw.w, (int)raw.z);
            atomicMax(mp.my_best + raw.x, pack_bid(bid, raw.y, pbits));
        }
    }
}

__global__ void __launch_bounds__(kWideThreads) mesh_max_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    const HotState h = load_hot(p.st);
    mesh_max_body(p, mp, h);
}

// =============================================================================================================
// K3: owner side, pass 2 -- winners install price and owner, reply bits go back to the bidders' ranks, evicted
// persons go to the ranks that hold them.
// =============================================================================================================
constexpr uint32_t kResolveChunk = 1024;     // entries one block resolves before it reserves room for their evictions

struct ResolveSmem {
    uint32_t evp[kResolveChunk], evw[kResolveChunk];   // evicted owners of a chunk: person, destination << 24 | rank
    uint32_t cnt[kMeshMaxRanks], base[kMeshMaxRanks], n;
};

__device__ __forceinline__ void mesh_resolve_body(const Params& p, const MeshParams& mp, const HotState& h, ResolveSmem& rs) {
    DevState* st = p.st;
    if (h.done) return;
    if (((volatile DevState*)st)->mesh_error) return;
    const uint32_t pbits = h.pbits;
    const bool nobody_owns = h.zero_prices != 0;       // round 1 (K4 clears the flag at its end)
    if (blockIdx.x == 0 && threadIdx.x == 0) mesh_stamp(mp, st, 8);
    const int lane = threadIdx.x & 31;
    const uint32_t warp = threadIdx.x >> 5;
    constexpr uint32_t kWarps = kWideThreads / 32;

    // evicted owners of a chunk are staged in `rs`: ONE global atomic per destination rank and chunk reserves their room
    uint32_t* const s_evp = rs.evp;
    uint32_t* const s_evw = rs.evw;
    uint32_t* const s_cnt = rs.cnt;
    uint32_t* const s_base = rs.base;
    uint32_t& s_n = rs.n;

    // The entries of all senders as ONE index space, every sender's region padded to a multiple of 32 (a warp's 32
    // consecutive indices then belong to one sender and make up one reply word); blocks take chunks of kResolveChunk
    // padded indices -- a short round is one chunk whatever the number of senders.
    uint32_t cnt[kMeshMaxRanks], poff[kMeshMaxRanks + 1];
    poff[0] = 0u;
#pragma unroll
    for (int a = 0; a < kMeshMaxRanks; ++a) {
        cnt[a] = ((uint32_t)a < mp.world) ? __ldcg(&mp.box[mp.rank]->bid_count[a]) : 0u;
        poff[a + 1] = poff[a] + ((cnt[a] + 31u) & ~31u);
    }
    const uint32_t total_padded = poff[kMeshMaxRanks];
    const uint32_t total_chunks = (total_padded + kResolveChunk - 1u) / kResolveChunk;
    const uint32_t working = mesh_working_blocks(total_chunks);
    if (blockIdx.x >= working) return;

    for (uint32_t c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        const uint32_t p0 = c * kResolveChunk, p1 = (p0 + kResolveChunk < total_padded) ? p0 + kResolveChunk : total_padded;
        if (threadIdx.x < kMeshMaxRanks) s_cnt[threadIdx.x] = 0u;
        if (threadIdx.x == 0) s_n = 0u;
        __syncthreads();
        // four entries per thread and pass, the loads of each stage issued together: the cell array of a large instance
        // (hundreds of MB) is far beyond the TLB's reach, a dependent random load costs microseconds, and only
        // memory-level parallelism hides it
        constexpr int UE = 4;
        for (uint32_t w0 = 0; w0 * 32u < p1 - p0; w0 += kWarps * UE) {
            uint4 raw[UE];
            unsigned long long word[UE];
            uint32_t prev[UE], snd[UE], eloc[UE];
            bool in[UE], won[UE];
#pragma unroll
            for (int u = 0; u < UE; ++u) {
                const uint32_t pi = p0 + (w0 + (uint32_t)u * kWarps + warp) * 32u + (uint32_t)lane;   // padded index
                uint32_t a = 0;
#pragma unroll
                for (int k = 1; k < kMeshMaxRanks; ++k) a += (pi >= poff[k]) ? 1u : 0u;
                uint32_t base_a = 0, cnt_a = 0;
#pragma unroll
                for (int k = 0; k < kMeshMaxRanks; ++k) { base_a = (a == (uint32_t)k) ? poff[k] : base_a; cnt_a = (a == (uint32_t)k) ? cnt[k] : cnt_a; }
                snd[u] = a;
                eloc[u] = pi - base_a;
                in[u] = pi < p1 && eloc[u] < cnt_a;
                raw[u] = make_uint4(0u, 0u, 0u, 0u);
                if (in[u]) raw[u] = __ldcg(reinterpret_cast<const uint4*>(mp.my_bid_in + (size_t)a * mp.cap_bid + eloc[u]));   // stored by another GPU: past the L1
            }
#pragma unroll
            for (int u = 0; u < UE; ++u) {
                word[u] = 0ull;
                if (in[u]) word[u] = __ldcg(mp.my_best + raw[u].x);
            }
#pragma unroll
            for (int u = 0; u < UE; ++u) {
                const double bid = __hiloint2double((int)raw[u].w, (int)raw[u].z);
                won[u] = in[u] && word[u] == pack_bid(bid, raw[u].y, pbits);
                prev[u] = SLA_DEV_NONE;
                if (won[u] && !nobody_owns) prev[u] = __ldcg(&(mp.my_cells + raw[u].x)->owner);   // first round: every object is free
            }
#pragma unroll
            for (int u = 0; u < UE; ++u) {
                if (won[u]) {
                    // one 128-bit store: {price = the winner's exact bid, owner = its person}
                    *reinterpret_cast<uint4*>(mp.my_cells + raw[u].x) = make_uint4(raw[u].z, raw[u].w, raw[u].y, 0u);
                    mp.my_best[raw[u].x] = 0ull;   // losers that look later see 0 or this word: neither equals theirs
                }
                const uint32_t wonmask = __ballot_sync(0xffffffffu, won[u]);
                // one word per 32 entries, back to the bidder's rank (the warp's 32 indices belong to one sender; words that
                // lie wholly in the padding are skipped)
                const uint32_t any_in = __ballot_sync(0xffffffffu, in[u]);
                if (lane == 0 && any_in) mp.reply_out[snd[u]][eloc[u] >> 5] = wonmask;
                const bool ev = prev[u] != SLA_DEV_NONE;
                const uint32_t evmask = __ballot_sync(0xffffffffu, ev);
                if (evmask) {
                    uint32_t slot0 = 0;
                    if (lane == 0) slot0 = atomicAdd(&s_n, (uint32_t)__popc(evmask));
                    slot0 = __shfl_sync(0xffffffffu, slot0, 0);
                    if (ev) {
                        const uint32_t dest = mesh_person_rank(mp, prev[u]);
                        const uint32_t rk = atomicAdd(&s_cnt[dest], 1u);
                        const uint32_t i = slot0 + (uint32_t)__popc(evmask & ((1u << lane) - 1u));
                        s_evp[i] = prev[u];
                        s_evw[i] = (dest << 24) | rk;
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < mp.world) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&mp.ev_cnt[threadIdx.x], s_cnt[threadIdx.x]) : 0u;
        __syncthreads();
        const uint32_t n = s_n;
        for (uint32_t i = threadIdx.x; i < n; i += kWideThreads) {
            const uint32_t w = s_evw[i], dest = w >> 24;
            mp.evict_out[dest][s_base[dest] + (w & 0xFFFFFFu)] = s_evp[i];     // to the rank that holds the person
        }
        __syncthreads();
    }

    if (mesh_last_block(&mp.tickets[2], true, working)) {
        const uint32_t epoch = ((volatile DevState*)st)->mesh_epoch;
        if (threadIdx.x < mp.world) {
            const uint32_t b = threadIdx.x;
            const uint32_t c = *reinterpret_cast<volatile uint32_t*>(&mp.ev_cnt[b]);
            *reinterpret_cast<volatile uint32_t*>(&mp.box[b]->evict_count[mp.rank]) = c;
            *reinterpret_cast<volatile uint32_t*>(&mp.ev_cnt[b]) = 0u;
        }
        if (threadIdx.x == 0) mesh_stamp(mp, st, 4);
        mesh_barrier_tail(mp, st, epoch + 2u);
        if (threadIdx.x == 0) mesh_stamp(mp, st, 5);
    }
}

__global__ void __launch_bounds__(kWideThreads) mesh_resolve_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    __shared__ ResolveSmem rs;
    const HotState h = load_hot(p.st);
    mesh_resolve_body(p, mp, h, rs);
}

// =============================================================================================================
// K4: bidder side -- outcomes of this rank's bids, intake of the evicted persons, next queue, round accounting.
// =============================================================================================================
struct FinishSmem {
    uint32_t emit[kAssignChunk];
    uint32_t cnt, base;
};

template <bool PERSIST>
__device__ __forceinline__ void mesh_finish_body(const Params& p, const MeshParams& mp, const HotState& h, FinishSmem& fs) {
    DevState* st = p.st;
    if (h.done) return;
    const uint32_t epoch = ((volatile DevState*)st)->mesh_epoch;
    if (mp.wait_at_start && !mesh_wait(mp, st, epoch + 2u)) return;
    if (((volatile DevState*)st)->mesh_error) return;
    const uint32_t cur = h.cur, qlen = h.qlen[cur & 1u];
    const bool identity = h.identity != 0;
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    uint32_t* __restrict__ next_queue = cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[(cur ^ 1u) & 1u];
    const uint32_t my_first = mp.row_begin[mp.rank];

    uint32_t* const s_emit = fs.emit;
    uint32_t& s_cnt = fs.cnt;
    uint32_t& s_base = fs.base;
    const int lane = threadIdx.x & 31;
    // work items: [0, qlen) the slots of this round's bidders, then the evict inbox regions rank by rank; a block takes
    // contiguous chunks of up to kAssignChunk items and reserves room in the next queue once per chunk
    uint32_t ev_cnt[kMeshMaxRanks], ev_total = 0;
    for (uint32_t a = 0; a < kMeshMaxRanks; ++a) {
        ev_cnt[a] = (a < mp.world) ? ld_cv_u32(&mp.box[mp.rank]->evict_count[a]) : 0u;
        ev_total += ev_cnt[a];
    }
    const uint32_t items = qlen + ev_total;
    uint32_t per = (items + gridDim.x - 1) / gridDim.x;
    per = ((per + kWideThreads - 1) / kWideThreads) * kWideThreads;
    if (per > (uint32_t)kAssignChunk) per = kAssignChunk;
    if (per == 0u) per = kWideThreads;
    const uint32_t working = mesh_working_blocks((items + per - 1u) / per);
    if (blockIdx.x >= working) return;
    for (uint32_t start = blockIdx.x * per; start < items; start += gridDim.x * per) {
        const uint32_t stop = (start + per < items) ? start + per : items;
        if (threadIdx.x == 0) s_cnt = 0u;
        __syncthreads();
        for (uint32_t t0 = start; t0 < stop; t0 += kWideThreads) {
            const uint32_t t = t0 + threadIdx.x;
            uint32_t emit = SLA_DEV_NONE;
            if (t < stop && t < qlen) {
                const uint32_t j = PERSIST ? __ldcg(p.slot_obj + t) : p.slot_obj[t];
                if (j != SLA_DEV_NONE) {
                    const uint32_t i = identity ? t : (PERSIST ? __ldcg(queue + t) : __ldg(queue + t));
                    const uint32_t pos = PERSIST ? __ldcg(mp.slot_pos + t) : mp.slot_pos[t];
                    bool won = false;
                    if (pos != SLA_DEV_NONE) {
                        const uint32_t g = j >> mp.view.shift;
                        const uint32_t word = ld_cv_u32(mp.my_reply_in + (size_t)g * mp.cap_words + (pos >> 5));
                        won = ((word >> (pos & 31u)) & 1u) != 0u;
                    }
                    if (won) p.p2o[i] = j; else emit = i;
                }
            } else if (t < stop) {
                uint32_t e = t - qlen, a = 0;
#pragma unroll
                for (int g = 0; g < kMeshMaxRanks - 1; ++g) {
                    const bool past = (a == (uint32_t)g) && (e >= ev_cnt[g]);
                    e -= past ? ev_cnt[g] : 0u;
                    a += past ? 1u : 0u;
                }
                const uint32_t person = ld_cv_u32(mp.my_evict_in + (size_t)a * mp.cap_evict + e) - my_first;
                p.p2o[person] = SLA_DEV_NONE;
                emit = person;
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
            if (ballot) {
                uint32_t wbase = 0;
                if (lane == 0) wbase = atomicAdd(&s_cnt, (uint32_t)__popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (emit != SLA_DEV_NONE) s_emit[wbase + __popc(ballot & ((1u << lane) - 1u))] = emit;
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = s_cnt ? atomicAdd(next_len, s_cnt) : 0u;
        __syncthreads();
        const uint32_t cnt = s_cnt, gbase = s_base;
        for (uint32_t e = threadIdx.x; e < cnt; e += kWideThreads) next_queue[gbase + e] = s_emit[e];
        __syncthreads();
    }

    if (mesh_last_block(&mp.tickets[3], false, working)) {
      if (threadIdx.x == 0) {
        mesh_stamp(mp, st, 6);
        volatile DevState* v = st;
        const uint32_t c = v->cur & 1u, ql = v->qlen[c];
        v->rounds = v->rounds + 1;
        v->wide_rounds = v->wide_rounds + 1;
        v->bids = v->bids + ql;
        if (v->regular_k) v->bid_arcs = v->bid_arcs + (unsigned long long)ql * v->regular_k;
        v->qlen[c] = 0u;
        v->cur = c ^ 1u;
        v->identity = 0u;
        v->zero_prices = 0u;
        const unsigned long long next = v->qlen[c ^ 1u];
        if (v->safety_rounds_left <= 1) { v->mesh_error = 2u; v->done = 1u; }
        else v->safety_rounds_left = v->safety_rounds_left - 1;
        for (uint32_t g = 0; g < mp.world; ++g)
            *reinterpret_cast<volatile unsigned long long*>(&mp.box[g]->next_total[mp.rank]) = next;
        mesh_stamp(mp, st, 7);                   // (stamped under the round that just ended)
        v->mesh_round = v->mesh_round + 1u;
        v->mesh_epoch = epoch + 2u;              // two barriers per round; the queue length travels under the next B1
        __threadfence_system();
      }
    }
}

__global__ void __launch_bounds__(kWideThreads) mesh_finish_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    __shared__ FinishSmem fs;
    const HotState h = load_hot(p.st);
    mesh_finish_body<false>(p, mp, h, fs);
}

// =============================================================================================================
// Tail engine: ONE persistent block per rank runs whole rounds -- the same four phases, a __syncthreads() where the
// grid-wide path has a kernel boundary -- once the ranks have at most kMeshTailPerRank bidders each on average (DevState::
// mesh_tail, set by the max kernel on every rank in the same round).  A short round on the grid-wide path is four
// launches of mostly idle grids plus two flag barriers, ~75 us at 8 GPUs; in here it is the two barriers and a few
// dependent memory round trips.  Everything another rank or an earlier round of this launch may have changed is read
// past the L1 (queue, prices, inboxes, the control block).
// LPR8 > 0: uniform-degree CSR with that many lanes per row (NARROW: u16 value mirror); LPR8 == 0: ragged CSR, LPR lanes.
// =============================================================================================================
template <int LPR8, int LPR, bool NARROW>
__global__ void __launch_bounds__(kWideThreads, 1) mesh_tail_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    union TailSmem {
        MeshStage stage;
        ResolveSmem rs;
        FinishSmem fs;
    };
    __shared__ TailSmem sm;
    DevState* st = p.st;
    {
        const volatile DevState* v = st;
        if (v->done || !v->mesh_tail || v->mesh_error) return;
    }
    for (;;) {
        HotState h = load_hot_cg(st);
        if (h.done || ((volatile DevState*)st)->mesh_error) break;
        if (LPR8 > 0) mesh_bid_body<(LPR8 > 0 ? LPR8 : 1), false, NARROW, true>(p, mp, h, sm.stage);
        else mesh_bid_ragged_body<(LPR > 0 ? LPR : 2), true>(p, mp, h, sm.stage);
        __syncthreads();
        h = load_hot_cg(st);
        mesh_max_body(p, mp, h);            // (also the termination test)
        __threadfence();
        __syncthreads();
        h = load_hot_cg(st);
        if (h.done || ((volatile DevState*)st)->mesh_error) break;
        mesh_resolve_body(p, mp, h, sm.rs);
        __syncthreads();
        mesh_finish_body<true>(p, mp, h, sm.fs);
        __threadfence();
        __syncthreads();
    }
}

// Initialisation of this rank's objects and persons (solver.rs:218-229) and the export of the solved cells into the
// plain prices / object_to_person arrays the C ABI hands out.
__global__ void __launch_bounds__(kWideThreads) mesh_init_kernel(const Params p, const __grid_constant__ MeshParams mp,
                                                                const uint32_t n_local) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t j = tid; j < mp.my_objects; j += stride) {
        ObjCell c;
        c.price = 0.0; c.owner = SLA_DEV_NONE; c.pad = 0u;
        mp.my_cells[j] = c;
        mp.my_best[j] = 0ull;
    }
    for (uint32_t i = tid; i < n_local; i += stride) p.p2o[i] = SLA_DEV_NONE;
    if (tid < mp.world) { mp.out_cnt[tid] = 0u; mp.ev_cnt[tid] = 0u; }
    if (tid < 4) mp.tickets[tid] = 0u;
}

__global__ void __launch_bounds__(kWideThreads) mesh_export_kernel(const Params p, const __grid_constant__ MeshParams mp) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t j = tid; j < mp.my_objects; j += stride) {
        p.prices[j] = mp.my_cells[j].price;
        p.o2p[j] = mp.my_cells[j].owner;
    }
}

}  // namespace sla

// =============================================================================================================
// Host side
// =============================================================================================================
struct sla_mesh_state {
    int rank = 0, world = 1;
    uint32_t row_begin[sla::kMeshMaxRanks + 1] = {};
    uint32_t global_rows = 0, global_cols = 0;
    uint32_t shift = 0, shard = 0;               // objects per rank = 2^shift
    uint32_t cap_bid = 0, cap_words = 0, cap_evict = 0;
    // layout of every rank's block (identical on all ranks): byte offsets
    size_t off_box = 0, off_cells = 0, off_best = 0, off_bid = 0, off_reply = 0, off_evict = 0, block_bytes = 0;
    unsigned char* block = nullptr;              // this rank's block (cudaMalloc: one IPC handle covers it)
    unsigned char* peer[sla::kMeshMaxRanks] = {};    // every rank's block as mapped here (own included)
    bool connected = false, active = false;
    uint32_t* d_slot_pos = nullptr;
    uint32_t* d_counters = nullptr;              // out_cnt[8] | ev_cnt[8] | tickets[4]
    unsigned long long* d_timeline = nullptr;    // SLA_MESH_TIMELINE=1: per-round globaltimer stamps of the kernels
    size_t cap_slot_pos = 0;
    sla::MeshParams mp;
    uint32_t epoch = 0;                          // barrier epoch at the start of the next solve (identical on all ranks)
    uint32_t learned_rounds = 0;                 // rounds + 1 of the previous solve of the resident shard (graph length)
    cudaGraphExec_t exec = nullptr;
    int exec_rounds = 0;
    uint64_t exec_generation = 0;
    const void* exec_vals16 = nullptr;
    bool exec_first = false;
    int exec_tail = -1;
    double eps = 0.0, gfirst = 0.0;
    int flip = 0;
    uint32_t launches = 0, graph_launches = 0;
    double timeout_s = 20.0;
};

namespace {

uint32_t mesh_lpr8(const sla_ctx* c) {           // lanes per row of the uniform-degree mesh scan: 1, 2 or 4
    uint32_t l = 1;
    while (l < 4 && l * 8u < c->regular_k) l *= 2;
    return l;
}

// which: 0 bid, 1 max, 2 resolve, 3 finish, 4 the persistent tail engine (concurrent ranks only)
void mesh_launch_phase(sla_ctx* c, const Params& p, int which, bool first_round, bool lockstep) {
    sla_mesh_state* ms = c->mesh;
    sla::MeshParams mp = ms->mp;
    mp.wait_at_start = lockstep ? 1u : 0u;
    mp.tail_engine = (!lockstep && c->opt_mesh_tail) ? 1u : 0u;
    if (which == 4) {
        const bool narrow = narrow_scan_ptr(c) != nullptr && p.vals16 != nullptr;
        if (use_regular(c)) {
            const uint32_t l = mesh_lpr8(c);
#define SLA_MESH_TAIL(L) do { if (narrow) mesh_tail_kernel<L, 0, true><<<1, kWideThreads, 0, c->stream>>>(p, mp); \
                              else mesh_tail_kernel<L, 0, false><<<1, kWideThreads, 0, c->stream>>>(p, mp); } while (0)
            if (l == 1) SLA_MESH_TAIL(1); else if (l == 2) SLA_MESH_TAIL(2); else SLA_MESH_TAIL(4);
#undef SLA_MESH_TAIL
        } else {
            if (c->lpr <= 2) mesh_tail_kernel<0, 2, false><<<1, kWideThreads, 0, c->stream>>>(p, mp);
            else if (c->lpr <= 8) mesh_tail_kernel<0, 8, false><<<1, kWideThreads, 0, c->stream>>>(p, mp);
            else mesh_tail_kernel<0, 32, false><<<1, kWideThreads, 0, c->stream>>>(p, mp);
        }
        return;
    }
    const int grid_bid = c->num_sms * 3;
    switch (which) {
        case 0:
            if (use_regular(c)) {
                // round 1 never reads prices (they are exactly 0 after init_solve, solver.rs:218-219) -- and it must not:
                // no barrier separates this rank's first scan from the other ranks' initialisation of their cells
                const bool zero = first_round;
                const bool narrow = narrow_scan_ptr(c) != nullptr && p.vals16 != nullptr;
                const uint32_t l = mesh_lpr8(c);
#define SLA_MESH_BID(L, Z, N) mesh_bid_kernel<L, Z, N><<<grid_bid, kWideThreads, 0, c->stream>>>(p, mp)
#define SLA_MESH_BID_L(L)                                                                                           \
    do {                                                                                                            \
        if (zero) { if (narrow) SLA_MESH_BID(L, true, true); else SLA_MESH_BID(L, true, false); }                   \
        else      { if (narrow) SLA_MESH_BID(L, false, true); else SLA_MESH_BID(L, false, false); }                 \
    } while (0)
                if (l == 1) SLA_MESH_BID_L(1); else if (l == 2) SLA_MESH_BID_L(2); else SLA_MESH_BID_L(4);
#undef SLA_MESH_BID_L
#undef SLA_MESH_BID
            } else {
                if (c->lpr <= 2) mesh_bid_ragged_kernel<2><<<grid_bid, kWideThreads, 0, c->stream>>>(p, mp);
                else if (c->lpr <= 8) mesh_bid_ragged_kernel<8><<<grid_bid, kWideThreads, 0, c->stream>>>(p, mp);
                else mesh_bid_ragged_kernel<32><<<grid_bid, kWideThreads, 0, c->stream>>>(p, mp);
            }
            break;
        // (four blocks per SM: every block fences at system scope and takes a ticket at its end -- more blocks only
        //  lengthen the short rounds)
        case 1: mesh_max_kernel<<<c->num_sms * 4, kWideThreads, 0, c->stream>>>(p, mp); break;
        case 2: mesh_resolve_kernel<<<c->num_sms * 4, kWideThreads, 0, c->stream>>>(p, mp); break;
        default: mesh_finish_kernel<<<c->num_sms * 4, kWideThreads, 0, c->stream>>>(p, mp); break;
    }
}

int mesh_check(sla_ctx* ctx, bool need_connected, bool need_active) {
    if (!ctx->mesh) return fail(ctx, SLA_ERR_STATE, "sla_mesh_create has not been called");
    if (need_connected && !ctx->mesh->connected) return fail(ctx, SLA_ERR_STATE, "sla_mesh_connect has not been called");
    if (need_active && !ctx->mesh->active) return fail(ctx, SLA_ERR_STATE, "sla_mesh_begin has not been called");
    return SLA_OK;
}

}  // namespace

extern "C" {

void sla_mesh_free(sla_ctx* ctx) {
    if (!ctx || !ctx->mesh) return;
    sla_mesh_state* ms = ctx->mesh;
    if (ms->exec) cudaGraphExecDestroy(ms->exec);
    cudaFree(ms->block);
    cudaFree(ms->d_slot_pos);
    cudaFree(ms->d_counters);
    cudaFree(ms->d_timeline);
    delete ms;
    ctx->mesh = nullptr;
}

int sla_mesh_create(sla_ctx* ctx, int rank, int world, const uint32_t* row_begins, uint32_t global_cols, void** block,
                    size_t* block_bytes) {
    if (!ctx) return SLA_ERR_INVALID;
    if (!row_begins || world < 1 || world > sla::kMeshMaxRanks || rank < 0 || rank >= world)
        return fail(ctx, SLA_ERR_INVALID, "mesh: world must be in [1, 8] and rank in [0, world)");
    if (!ctx->has_csr) return fail(ctx, SLA_ERR_STATE, "sla_mesh_create called before this rank's CSR shard was uploaded");
    for (int g = 0; g < world; ++g)
        if (row_begins[g + 1] < row_begins[g]) return fail(ctx, SLA_ERR_INVALID, "mesh: row_begins must be non-decreasing");
    if (row_begins[0] != 0 || row_begins[rank + 1] - row_begins[rank] != ctx->n_rows)
        return fail(ctx, SLA_ERR_INVALID, "mesh: row_begins does not match the uploaded shard");
    if (global_cols != ctx->n_cols) return fail(ctx, SLA_ERR_INVALID, "mesh: the shard must be uploaded with the global number of columns");
    if (row_begins[world] > global_cols) return fail(ctx, SLA_ERR_INVALID, "num_rows must be <= num_cols");
    CU(cudaSetDevice(ctx->device));
    sla_mesh_free(ctx);
    sla_mesh_state* ms = new sla_mesh_state();
    ctx->mesh = ms;
    ms->rank = rank; ms->world = world;
    for (int g = 0; g <= world; ++g) ms->row_begin[g] = row_begins[g];
    ms->global_rows = row_begins[world];
    ms->global_cols = global_cols;
    // objects per rank: the power of two at or above ceil(M / world), so that owner and local index are a shift and a mask
    const uint32_t per = (global_cols + (uint32_t)world - 1u) / (uint32_t)world;
    uint32_t shift = 0;
    while (((uint64_t)1 << shift) < per) ++shift;
    ms->shift = shift;
    ms->shard = (uint32_t)1 << shift;
    uint32_t max_rows = 1;
    for (int g = 0; g < world; ++g) max_rows = std::max(max_rows, row_begins[g + 1] - row_begins[g]);
    ms->cap_bid = (max_rows + 31u) & ~31u;
    ms->cap_words = ms->cap_bid / 32u;
    ms->cap_evict = ms->cap_bid;
    auto align256 = [](size_t x) { return (x + 255) & ~(size_t)255; };
    size_t off = 0;
    ms->off_box = off;     off = align256(off + sizeof(sla::MeshMailbox));
    ms->off_cells = off;   off = align256(off + (size_t)ms->shard * sizeof(sla::ObjCell));
    ms->off_best = off;    off = align256(off + (size_t)ms->shard * sizeof(unsigned long long));
    ms->off_bid = off;     off = align256(off + (size_t)world * ms->cap_bid * sizeof(sla::BidEntry));
    ms->off_reply = off;   off = align256(off + (size_t)world * ms->cap_words * sizeof(uint32_t));
    ms->off_evict = off;   off = align256(off + (size_t)world * ms->cap_evict * sizeof(uint32_t));
    ms->block_bytes = off;
    cudaError_t e = cudaMalloc((void**)&ms->block, ms->block_bytes);
    if (e != cudaSuccess) { ms->block = nullptr; return fail(ctx, SLA_ERR_ALLOC, std::string("cudaMalloc(mesh block): ") + cudaGetErrorString(e)); }
    CU(cudaMemsetAsync(ms->block, 0, sizeof(sla::MeshMailbox), ctx->stream));
    int rc;
    if ((rc = dev_alloc(ctx, &ms->d_slot_pos, (size_t)ctx->cap_rows))) return rc;
    ms->cap_slot_pos = ctx->cap_rows;
    if ((rc = dev_alloc(ctx, &ms->d_counters, 32))) return rc;
    CU(cudaMemsetAsync(ms->d_counters, 0, 32 * sizeof(uint32_t), ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (const char* t = getenv("SLA_MESH_TIMEOUT_S")) ms->timeout_s = atof(t);
    if (const char* t = getenv("SLA_MESH_TIMELINE")) {
        if (atoi(t) != 0) {
            if ((rc = dev_alloc(ctx, &ms->d_timeline, (size_t)sla::kMeshTimelineRounds * sla::kMeshTimelineSlots))) return rc;
            CU(cudaMemset(ms->d_timeline, 0, (size_t)sla::kMeshTimelineRounds * sla::kMeshTimelineSlots * sizeof(unsigned long long)));
        }
    }
    if (block) *block = ms->block;
    if (block_bytes) *block_bytes = ms->block_bytes;
    return SLA_OK;
}

// Peer mappings across processes: the exporting rank's cudaMalloc'ed block as a 64-byte handle, opened by the others.
int sla_ipc_export(void* dev_ptr, unsigned char* handle64) {
    if (!dev_ptr || !handle64) return SLA_ERR_INVALID;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries IPC handles as 64 bytes");
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, dev_ptr) != cudaSuccess) { cudaGetLastError(); return SLA_ERR_CUDA; }
    memcpy(handle64, &h, 64);
    return SLA_OK;
}
int sla_ipc_import(int device, const unsigned char* handle64, void** dev_ptr) {
    if (!handle64 || !dev_ptr) return SLA_ERR_INVALID;
    *dev_ptr = nullptr;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return SLA_ERR_CUDA; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    cudaError_t e = cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e);
        cudaGetLastError();
        *dev_ptr = nullptr;
        return SLA_ERR_CUDA;
    }
    return SLA_OK;
}
int sla_ipc_release(int device, void* dev_ptr) {
    if (!dev_ptr) return SLA_OK;
    cudaSetDevice(device);
    cudaError_t e = cudaIpcCloseMemHandle(dev_ptr);
    cudaGetLastError();
    return e == cudaSuccess ? SLA_OK : SLA_ERR_CUDA;
}

// peer_blocks[g]: rank g's block as addressable from this context's device -- its own block for g == rank, the pointer
// sla_ipc_import returned (other process), or the other context's pointer (same process; peer access between the two
// devices is enabled here when they differ).  peer_devices may be NULL (all blocks live on this device or are IPC maps).
int sla_mesh_connect(sla_ctx* ctx, void* const* peer_blocks, const int* peer_devices) {
    if (!ctx || !peer_blocks) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, false, false);
    if (rc) return rc;
    sla_mesh_state* ms = ctx->mesh;
    CU(cudaSetDevice(ctx->device));
    for (int g = 0; g < ms->world; ++g) {
        if (!peer_blocks[g]) return fail(ctx, SLA_ERR_INVALID, "mesh: null peer block");
        if (g == ms->rank && peer_blocks[g] != ms->block) return fail(ctx, SLA_ERR_INVALID, "mesh: peer_blocks[rank] must be this rank's own block");
        ms->peer[g] = static_cast<unsigned char*>(peer_blocks[g]);
        if (peer_devices && peer_devices[g] >= 0 && peer_devices[g] != ctx->device) {
            int can = 0;
            cudaDeviceCanAccessPeer(&can, ctx->device, peer_devices[g]);
            if (!can) return fail(ctx, SLA_ERR_CUDA, "mesh: no peer access between device " + std::to_string(ctx->device) + " and " + std::to_string(peer_devices[g]));
            cudaError_t e = cudaDeviceEnablePeerAccess(peer_devices[g], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, SLA_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
            cudaGetLastError();
        }
    }
    sla::MeshParams& mp = ms->mp;
    memset(&mp, 0, sizeof mp);
    mp.view.shift = ms->shift;
    mp.view.mask = ms->shard - 1u;
    for (int g = 0; g < ms->world; ++g) {
        unsigned char* b = ms->peer[g];
        mp.view.cells[g] = reinterpret_cast<const sla::ObjCell*>(b + ms->off_cells);
        mp.box[g] = reinterpret_cast<sla::MeshMailbox*>(b + ms->off_box);
        mp.bid_out[g] = reinterpret_cast<sla::BidEntry*>(b + ms->off_bid) + (size_t)ms->rank * ms->cap_bid;
        mp.reply_out[g] = reinterpret_cast<uint32_t*>(b + ms->off_reply) + (size_t)ms->rank * ms->cap_words;
        mp.evict_out[g] = reinterpret_cast<uint32_t*>(b + ms->off_evict) + (size_t)ms->rank * ms->cap_evict;
    }
    mp.my_cells = reinterpret_cast<sla::ObjCell*>(ms->block + ms->off_cells);
    mp.my_best = reinterpret_cast<unsigned long long*>(ms->block + ms->off_best);
    mp.my_bid_in = reinterpret_cast<sla::BidEntry*>(ms->block + ms->off_bid);
    mp.my_reply_in = reinterpret_cast<uint32_t*>(ms->block + ms->off_reply);
    mp.my_evict_in = reinterpret_cast<uint32_t*>(ms->block + ms->off_evict);
    mp.slot_pos = ms->d_slot_pos;
    mp.out_cnt = ms->d_counters;
    mp.ev_cnt = ms->d_counters + 8;
    mp.tickets = ms->d_counters + 16;
    mp.rank = (uint32_t)ms->rank;
    mp.world = (uint32_t)ms->world;
    mp.cap_bid = ms->cap_bid; mp.cap_words = ms->cap_words; mp.cap_evict = ms->cap_evict;
    const uint64_t first_obj = (uint64_t)ms->rank << ms->shift;
    mp.my_objects = first_obj >= ms->global_cols ? 0u : (uint32_t)std::min<uint64_t>(ms->shard, ms->global_cols - first_obj);
    for (int g = 0; g <= ms->world; ++g) mp.row_begin[g] = ms->row_begin[g];
    for (int g = ms->world + 1; g <= sla::kMeshMaxRanks; ++g) mp.row_begin[g] = ms->row_begin[ms->world];
    mp.timeout_ns = (unsigned long long)(ms->timeout_s * 1e9);
    mp.timeline = ms->d_timeline;
    ms->connected = true;
    if (ms->exec) { cudaGraphExecDestroy(ms->exec); ms->exec = nullptr; }
    return SLA_OK;
}

// Start of a solve (KhoslaSolver semantics, ksparse.rs:153-184, for the GLOBAL instance): global_w_min / global_w_max /
// global_first_value as for sla_part_begin.  No inter-rank synchronisation is needed before or after this call.
int sla_mesh_begin(sla_ctx* ctx, int maximize, double eps, double global_w_min, double global_w_max, double global_first_value) {
    if (!ctx) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, false);
    if (rc) return rc;
    if (!ctx->has_csr) return fail(ctx, SLA_ERR_STATE, "sla_mesh_begin called before the CSR shard was uploaded");
    sla_mesh_state* ms = ctx->mesh;
    if (ctx->n_rows != ms->row_begin[ms->rank + 1] - ms->row_begin[ms->rank] || ctx->n_cols != ms->global_cols)
        return fail(ctx, SLA_ERR_STATE, "mesh: the resident shard no longer matches sla_mesh_create");
    CU(cudaSetDevice(ctx->device));
    if (ctx->cap_rows > ms->cap_slot_pos) {
        if ((rc = dev_alloc(ctx, &ms->d_slot_pos, (size_t)ctx->cap_rows))) return rc;
        ms->cap_slot_pos = ctx->cap_rows;
        ms->mp.slot_pos = ms->d_slot_pos;
        if (ms->exec) { cudaGraphExecDestroy(ms->exec); ms->exec = nullptr; }
    }
    const uint32_t N = ctx->n_rows, M = ctx->n_cols;
    const bool flip = (maximize != 0) != (global_first_value >= 0.0);      // solver.rs:207-216, decided by the first value
    ctx->dev_sign = flip ? -1 : 1;
    const double w_min = flip ? -global_w_max : global_w_min, w_max = flip ? -global_w_min : global_w_max;
    DevState s;
    memset(&s, 0, sizeof s);
    s.qlen[0] = N;
    s.identity = 1;
    s.zero_prices = 1;
    s.algo = ALGO_KHOSLA;
    s.pbits = person_bits(ms->global_rows);
    s.tail_max = 0;
    s.skip_zero = 1u;                      // see mesh_launch_phase: round 1 must not read other ranks' cells
    s.sign_flip = flip ? 0x80000000u : 0u;
    s.n_rows = N;
    s.n_cols = M;
    s.person_base = ms->row_begin[ms->rank];
    s.regular_k = use_regular(ctx) ? ctx->regular_k : 0u;
    s.safety_rounds_left = 1ull << 40;
    s.max_iterations = 0xFFFFFFFFu;
    const double m = (double)M;
    s.eps = std::isnan(eps) ? 1.0 / m : eps;                        // ksparse.rs:162-169
    s.target_eps = s.eps;
    s.threshold = (m / 2.0) * (w_max - w_min + s.eps);               // ksparse.rs:181
    s.prune_ok = (ctx->opt_prune && s.eps >= 0.0) ? 1u : 0u;
    s.mesh_round = 1;
    s.mesh_epoch = ms->epoch;
    *ctx->h_state = s;
    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_state, ctx->h_state, sizeof(DevState), cudaMemcpyHostToDevice, ctx->stream));
    const Params p = make_params(ctx);
    mesh_init_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ms->mp, N);
    CU(cudaGetLastError());
    ms->active = true;
    ms->eps = s.eps;
    ms->flip = flip ? 1 : 0;
    ms->gfirst = global_first_value;
    ms->launches = 1;
    ms->graph_launches = 0;
    ctx->has_solution = false;
    ctx->best_dirty = true;
    return SLA_OK;
}

// One kernel of the current round, launched eagerly and waited for (which: 0 bid, 1 max, 2 resolve, 3 finish).  For
// lockstep operation -- several ranks driven by one host thread, e.g. all of them on ONE GPU in the tests, where a
// kernel that waits for another rank's flag must never be resident before that rank's kernel has run: the caller runs
// phase k on every rank before phase k + 1 on any.
int sla_mesh_phase(sla_ctx* ctx, int which) {
    if (!ctx) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, true);
    if (rc) return rc;
    if (which < 0 || which > 3) return fail(ctx, SLA_ERR_INVALID, "mesh: phase must be in [0, 3]");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    // (round 1 is the only round whose state says zero_prices; the host mirrors it from the last poll)
    const bool first_round = ctx->h_state->mesh_round <= 1u && ctx->h_state->zero_prices != 0u;
    mesh_launch_phase(ctx, p, which, first_round, true);
    ctx->mesh->launches += 1;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    return SLA_OK;
}

// Reads the control block back: done (the solve has ended on every rank), the round about to run, this rank's queue.
int sla_mesh_poll(sla_ctx* ctx, int* done, uint32_t* round, uint32_t* local_queue_len) {
    if (!ctx) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, true);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    if ((rc = poll_state(ctx))) return rc;
    const DevState& f = *ctx->h_state;
    if (f.mesh_error == 1u) return fail(ctx, SLA_ERR_STATE, "mesh: a barrier gave up waiting for a peer rank (timeout)");
    if (f.mesh_error == 2u) return fail(ctx, SLA_ERR_STATE, "mesh: safety round limit reached");
    if (done) *done = f.done ? 1 : 0;
    if (round) *round = f.mesh_round;
    if (local_queue_len) *local_queue_len = f.qlen[f.cur & 1u];
    return SLA_OK;
}

// The whole solve: graphs of `R` rounds (4 kernels each, the first with the zero-price scan) until the device reports
// done.  Every rank calls this at (about) the same time; the ranks only meet in the kernels' flag barriers.
int sla_mesh_solve(sla_ctx* ctx) {
    if (!ctx) return SLA_ERR_INVALID;
    NvtxRange nvtx("sla_mesh_solve");
    int rc = mesh_check(ctx, true, true);
    if (rc) return rc;
    sla_mesh_state* ms = ctx->mesh;
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    const auto t_start = std::chrono::steady_clock::now();
    bool first = true, done = false;
    while (!done) {
        // graph of the first launch: round 1 (zero-price scan) + as many rounds as the previous solve took; later
        // launches: 8 rounds of the general shape
        const int rounds = first ? (ms->learned_rounds ? (int)std::min<uint32_t>(ms->learned_rounds, 256u) : 8) : 8;
        if (!ms->exec || ms->exec_rounds != rounds || ms->exec_generation != ctx->generation ||
            ms->exec_vals16 != narrow_scan_ptr(ctx) || ms->exec_first != first || ms->exec_tail != ctx->opt_mesh_tail) {
            if (ms->exec) { cudaGraphExecDestroy(ms->exec); ms->exec = nullptr; }
            cudaGraph_t graph = nullptr;
            CU(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
            const int phases = ctx->opt_mesh_tail ? 5 : 4;     // the tail engine sits behind every round's finish kernel
            for (int r = 0; r < rounds; ++r)
                for (int k = 0; k < phases; ++k) {
                    // the first round's scan + push is bracketed by two event-record nodes (sla_mesh_round1_ms)
                    const bool bracket = first && r == 0 && k == 0;
                    if (bracket) cudaEventRecordWithFlags(ctx->ev[3], ctx->stream, cudaEventRecordExternal);
                    mesh_launch_phase(ctx, p, k, first && r == 0, false);
                    if (bracket) cudaEventRecordWithFlags(ctx->ev[4], ctx->stream, cudaEventRecordExternal);
                }
            cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
            if (e != cudaSuccess) return fail(ctx, SLA_ERR_CUDA, std::string("cudaStreamEndCapture(mesh): ") + cudaGetErrorString(e));
            e = cudaGraphInstantiate(&ms->exec, graph, 0);
            cudaGraphDestroy(graph);
            if (e != cudaSuccess) { ms->exec = nullptr; return fail(ctx, SLA_ERR_CUDA, std::string("cudaGraphInstantiate(mesh): ") + cudaGetErrorString(e)); }
            ms->exec_rounds = rounds;
            ms->exec_generation = ctx->generation;
            ms->exec_vals16 = narrow_scan_ptr(ctx);
            ms->exec_first = first;
            ms->exec_tail = ctx->opt_mesh_tail;
        }
        CU(cudaGraphLaunch(ms->exec, ctx->stream));
        ms->graph_launches += 1;
        ms->launches += (uint32_t)((ctx->opt_mesh_tail ? 5 : 4) * rounds);
        first = false;
        int d = 0;
        if ((rc = sla_mesh_poll(ctx, &d, nullptr, nullptr))) return rc;
        done = d != 0;
        if (!done && std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count() > ctx->opt_timeout_s)
            return fail(ctx, SLA_ERR_STATE, "mesh solve exceeded the wall-clock guard (timeout_s)");
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    return SLA_OK;
}

// End of a solve: exports this rank's cells into plain arrays and (optionally) downloads this rank's part of the result:
// person_to_object for its rows (global object ids), object_to_person (global person ids) and prices for the objects it
// owns -- objects [rank << shift, ...), `num_owned` of them (sla_mesh_owned).
int sla_mesh_finish(sla_ctx* ctx, uint32_t* person_to_object, uint32_t* object_to_person, double* prices, sla_stats* stats) {
    if (!ctx) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, true);
    if (rc) return rc;
    sla_mesh_state* ms = ctx->mesh;
    CU(cudaSetDevice(ctx->device));
    join_workers(ctx);      // the in-place negation of the host `values` an upload may have started (solver.rs:214-216)
    if ((rc = poll_state(ctx))) return rc;
    const DevState f = *ctx->h_state;
    if (!f.done) return fail(ctx, SLA_ERR_STATE, "mesh: sla_mesh_finish called before the solve has ended");
    if (f.mesh_error) return fail(ctx, SLA_ERR_STATE, "mesh: the solve ended with an error (barrier timeout or safety limit)");
    const Params p = make_params(ctx);
    mesh_export_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ms->mp);
    const uint32_t owned = ms->mp.my_objects;
    if (person_to_object) CU(cudaMemcpyAsync(person_to_object, ctx->d_p2o, (size_t)ctx->n_rows * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (object_to_person && owned) CU(cudaMemcpyAsync(object_to_person, ctx->d_o2p, (size_t)owned * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (prices && owned) CU(cudaMemcpyAsync(prices, ctx->d_prices, (size_t)owned * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (stats) {
        memset(stats, 0, sizeof *stats);
        stats->num_unassigned = f.dropped;            // local share; the caller sums over ranks
        stats->nits = (uint32_t)f.bids;
        stats->eps = ms->eps;
        stats->rounds = f.rounds;
        stats->bids = f.bids;
        stats->bid_arcs = f.bid_arcs;
        stats->dropped = f.dropped;
        stats->values_negated = (uint32_t)ms->flip;
        stats->wide_rounds = f.wide_rounds;
        stats->kernel_launches = ms->launches + 1;
        stats->graph_launches = ms->graph_launches;
        if (ms->graph_launches) {
            cudaEventElapsedTime(&stats->ms_solve, ctx->ev[0], ctx->ev[1]);
            cudaEventElapsedTime(&stats->ms_total, ctx->ev[0], ctx->ev[2]);
        }
    }
    ms->epoch = f.mesh_epoch;
    // graph length of the next solve of this resident shard: up to the round in which the tail engine took over, or
    // all rounds plus the one that notices the end
    ms->learned_rounds = f.mesh_tail ? std::max<uint32_t>(f.mesh_switch_round, 1u)
                                     : (uint32_t)std::min<unsigned long long>(f.rounds + 1ull, 256ull);
    ms->active = false;
    ctx->has_solution = false;      // the plain arrays hold this rank's slices only: the single-GPU post-processing calls do not apply
    ctx->best_dirty = true;
    return SLA_OK;
}

// Duration of the first round's bid kernel (scan of all local rows + push of the bids) of the last solve, from the two
// event-record nodes around it in the first graph; this rank's objective share (sum of the chosen arcs' values of the
// local rows, exact for integer weights; get_objective, solver.rs:110-142) -- the caller adds the shares up.
int sla_mesh_round1_ms(sla_ctx* ctx, float* bid_ms) {
    if (!ctx || !bid_ms) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, false);
    if (rc) return rc;
    *bid_ms = 0.f;
    if (cudaEventElapsedTime(bid_ms, ctx->ev[3], ctx->ev[4]) != cudaSuccess) { cudaGetLastError(); *bid_ms = 0.f; }
    return SLA_OK;
}

int sla_mesh_objective(sla_ctx* ctx, double* objective) {
    if (!ctx || !objective) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, true, false);
    if (rc) return rc;
    if (ctx->mesh->active) return fail(ctx, SLA_ERR_STATE, "mesh: sla_mesh_objective needs a finished solve");
    CU(cudaSetDevice(ctx->device));
    const Params p = make_params(ctx);
    const uint32_t flip = ctx->dev_sign < 0 ? 0x80000000u : 0u;
    objective_kernel<<<ctx->grid_wide, kWideThreads, 0, ctx->stream>>>(p, ctx->n_rows, flip, ctx->d_partial);
    CU(cudaMemcpyAsync(ctx->h_partial, ctx->d_partial, (size_t)ctx->grid_wide * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    double sum = 0.0;
    for (int b = 0; b < ctx->grid_wide; ++b) sum += ctx->h_partial[b];
    // reference src/solver.rs:111-137: the sign is taken from the (current, possibly negated) first value of the instance
    const double first_eff = ctx->mesh->flip ? -ctx->mesh->gfirst : ctx->mesh->gfirst;
    *objective = (first_eff >= 0.0) ? sum : -sum;
    return SLA_OK;
}

// Development aid (SLA_MESH_TIMELINE=1 at sla_mesh_create): globaltimer stamps of the last solve, 12 slots per round
// (round r at out[12 r ..]): 0 bid kernel starts, 1 its last block is done locally, 2 barrier B1 passed, 3 max kernel
// starts, 8 resolve kernel starts, 4 resolve kernel done locally, 5 barrier B2 passed, 6 finish kernel done locally,
// 7 round accounted.
int sla_mesh_timeline(sla_ctx* ctx, unsigned long long* out, size_t capacity) {
    if (!ctx || !out) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, false, false);
    if (rc) return rc;
    if (!ctx->mesh->d_timeline) return fail(ctx, SLA_ERR_STATE, "mesh: no timeline (set SLA_MESH_TIMELINE=1 before sla_mesh_create)");
    const size_t n = std::min<size_t>(capacity, (size_t)sla::kMeshTimelineRounds * sla::kMeshTimelineSlots);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpy(out, ctx->mesh->d_timeline, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return SLA_OK;
}

// Facts about the partition: objects per rank (2^shift), how many of them this rank owns, its first object and first row.
int sla_mesh_owned(sla_ctx* ctx, uint32_t* shard_objects, uint32_t* num_owned, uint32_t* first_object, uint32_t* first_row) {
    if (!ctx) return SLA_ERR_INVALID;
    int rc = mesh_check(ctx, false, false);
    if (rc) return rc;
    sla_mesh_state* ms = ctx->mesh;
    const uint64_t first_obj = (uint64_t)ms->rank << ms->shift;
    if (shard_objects) *shard_objects = ms->shard;
    if (num_owned) *num_owned = first_obj >= ms->global_cols ? 0u : (uint32_t)std::min<uint64_t>(ms->shard, ms->global_cols - first_obj);
    if (first_object) *first_object = (uint32_t)std::min<uint64_t>(first_obj, ms->global_cols);
    if (first_row) *first_row = ms->row_begin[ms->rank];
    return SLA_OK;
}

}  // extern "C"
