// sla_common.cuh -- shared device-side definitions of the B200 auction path.
//
// Round semantics: a synchronous Jacobi auction.  Every queued (unassigned) person scans its CSR
// row against frozen prices (the reference's choice rule, src/ksparse.rs:199-214 == src/symmetric.rs:361-376),
// submits one bid, the per-object winner is the maximum packed (bid key, person) word, the winner installs its
// own exact f64 bid as the new price, and evicted owners plus losers form the next queue.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SLA_DEV_NONE 0xFFFFFFFFu

namespace sla {

constexpr int kTailCap = 1024;       // max bidders the single-CTA tail engine accepts (smem-resident queue)
constexpr int kWideThreads = 256;    // block size of the grid-wide kernels
#ifndef SLA_TAIL_THREADS
#define SLA_TAIL_THREADS 768
#endif
constexpr int kTailThreads = SLA_TAIL_THREADS;   // block size of the tail engine (24 warps: 80 registers per thread)
constexpr uint32_t kTailSlots = kTailThreads / 32;   // bidders of a "small" round: one warp per slot
// Cluster engine (mid_kernel): one thread-block cluster runs the rounds whose queue is too long for one CTA without a
// shared-memory price mirror and too short to be worth two grid-wide launches.
constexpr int kMidCtas = 8;                                   // portable cluster size
constexpr int kMidThreads = 1024;
constexpr uint32_t kMidLocalCap = 1024;                       // bidders one CTA of the cluster holds
constexpr uint32_t kMidMax = kMidCtas * kMidLocalCap;         // longest queue the cluster engine accepts

enum : uint32_t { ALGO_KHOSLA = 0, ALGO_FORWARD = 1 };
enum : uint32_t { ACTION_NONE = 0, ACTION_RESET = 1, ACTION_RESET_ALL = 2 };   // RESET: wipe the assignment; RESET_ALL: and the prices

// Device-resident control block of one solve.  Only single threads write it (see the kernels).
// The first 112 bytes are the fields every kernel needs at start-up; HotState mirrors them so that a kernel can
// fetch them with seven independent 128-bit loads (one L2 round trip instead of a chain of dependent ones).
struct DevState {
    // ---- hot (mirrored by HotState) ----
    uint32_t qlen[2];        // lengths of the two queue buffers
    uint32_t cur;            // which queue buffer holds the current bidders
    uint32_t done;           // solve finished
    uint32_t identity;       // current queue is the implicit [0, qlen)
    uint32_t zero_prices;    // every price is exactly 0.0 (first round after init_solve): gather may be skipped
    uint32_t algo;
    uint32_t pbits;          // bits of the person field of a packed bid word
    uint32_t tail_max;
    uint32_t skip_zero;      // option zero_price_skip
    uint32_t sign_flip;      // 0x80000000 when the effective values are the negated uploaded values, else 0
    uint32_t regular_k;      // every row has exactly this many arcs and it is a multiple of 8 (0 = ragged CSR)
    uint32_t n_rows, n_cols;
    uint32_t start_opt;      // start_from_optimal_eps (symmetric.rs:251-266)
    uint32_t action;
    double eps;
    double threshold;        // Khosla price threshold (ksparse.rs:181)
    double target_eps;
    double tol;
    uint32_t person_base;    // global id of local row 0 (row-partitioned instance; 0 otherwise)
    uint32_t own_mode;       // tail engine: owners mirrored in shared memory (0 no, 1 as u32, 2 as u16)
    uint32_t tail_cap;       // tail engine: capacity of its shared-memory queue arrays (power of two, >= tail_max)
    uint32_t hot_pad[1];
    // ---- cold ----
    uint32_t nits;           // Forward: rounds (symmetric.rs:277); Khosla: filled from `bids` at the end
    uint32_t nreductions;
    uint32_t optimal;
    uint32_t ecs_violated;
    uint32_t ecs_ticket;
    uint32_t dropped;
    uint32_t max_iterations;
    uint32_t tail_round_cap; // rounds one tail launch may run before handing control back to the host
    uint32_t kscale;         // Khosla rounds currently run under the eps-schedule (square instances)
    uint32_t assign_ticket;  // blocks of assign_wide_kernel that have finished (the last one runs control step A)
    uint32_t prune_ok;       // prices are known to be >= 0 for the whole solve (eps >= 0): profit <= value, so the
                             // gathering scan may skip arcs whose value is below a proven bound on the second-best profit
    uint32_t wide_ctl_done;  // control step A of this super-round already ran (in assign_wide_kernel's last block)
    // mesh engine (sla_mesh.cuh): round number of the current solve (from 1), barrier epoch at the start of the round
    // (keeps counting across solves), 1 = a barrier gave up / 2 = safety limit
    uint32_t mesh_round, mesh_epoch, mesh_error;
    uint32_t mesh_tail;      // the persistent one-block tail engine takes the rounds from here on (short rounds)
    uint32_t mesh_switch_round;   // round in which mesh_tail was set (graph length of the next solve)
    uint32_t tail_own_max;   // bidders at or below which the single-CTA tail engine runs; queues in (tail_own_max, tail_max]
                             // belong to the cluster engine when there is one (otherwise tail_own_max == tail_max)
    uint32_t cluster_rounds; // rounds the cluster engine ran (also counted in tail_rounds)
    uint32_t mesh_pad[1];
    unsigned long long rounds, bids, bid_arcs, wide_rounds, tail_rounds;
    unsigned long long safety_rounds_left;
    unsigned long long dbg[24];  // cycle counters of the tail engine when built with -DSLA_TAIL_TIMING
};

struct alignas(16) HotState {
    uint32_t qlen[2], cur, done;
    uint32_t identity, zero_prices, algo, pbits;
    uint32_t tail_max, skip_zero, sign_flip, regular_k;
    uint32_t n_rows, n_cols, start_opt, action;
    double eps, threshold;
    double target_eps, tol;
    uint32_t person_base, own_mode, tail_cap, hot_pad[1];
};
static_assert(sizeof(HotState) == 112, "HotState must mirror the first 112 bytes of DevState");
static_assert(sizeof(DevState) % 16 == 0, "DevState is copied as 128-bit words");

// The same past the L1 (persistent kernels re-read the block every round).
__device__ __forceinline__ HotState load_hot_cg(const DevState* st) {
    HotState h;
    const uint4* src = reinterpret_cast<const uint4*>(st);
    uint4* dst = reinterpret_cast<uint4*>(&h);
#pragma unroll
    for (int i = 0; i < 7; ++i) dst[i] = __ldcg(src + i);
    return h;
}

__device__ __forceinline__ HotState load_hot(const DevState* st) {
    HotState h;
    const uint4* src = reinterpret_cast<const uint4*>(st);
    uint4* dst = reinterpret_cast<uint4*>(&h);
#pragma unroll
    for (int i = 0; i < 7; ++i) dst[i] = src[i];   // L1-cached: thousands of warps read the same 96 bytes (the L1 is
                                                   // invalidated at every launch boundary, so the data is current)
    return h;
}

// Value statistics of the uploaded CSR (computed once per upload).
struct DevCsrStats {
    unsigned long long min_key, max_key;   // order-preserving u64 keys of min / max value
    unsigned long long bad_cols;           // arcs whose column index is >= num_cols
    unsigned long long bad_rows;           // rows whose extents are not monotone / exceed nnz
    unsigned long long irregular_rows;     // rows whose degree differs from row 0's
    unsigned long long not_u16;            // values that are not integers in [0, 65535] with a clear sign bit
};

// true when (double)(uint16_t)x reproduces x bit for bit
__device__ __forceinline__ bool is_u16_value(double x) {
    if (!(x >= 0.0 && x <= 65535.0)) return false;            // also rejects NaN
    const uint32_t q = (uint32_t)x;
    return __double_as_longlong((double)q) == __double_as_longlong(x);   // rejects fractions and -0.0
}

// Every buffer a kernel needs, passed by value.  Pointers only: sizes, sign and all per-solve scalars live in
// the device-resident DevState, so a captured graph stays valid across uploads and solves until a buffer is
// reallocated.
struct Params {
    const uint32_t* __restrict__ row_ptr;
    const uint32_t* __restrict__ cols;
    const double* __restrict__ vals;
    const uint16_t* __restrict__ vals16;   // lossless u16 mirror of `vals` left by a narrow upload (nullptr: none)
    double* prices;
    uint32_t* p2o;
    uint32_t* o2p;
    unsigned long long* best;   // packed best-bid word per object, 0 = no bid
    uint32_t* queue[2];
    uint32_t* slot_obj;         // per queue slot: object bid on (SLA_DEV_NONE = dropped)
    double* slot_bid;           // per queue slot: exact f64 bid
    DevState* st;
};

// Dynamic shared memory of the tail engine: [price mirror][owner mirror][hash words][hash keys][per-slot arrays].
// Shared by the kernel (carving) and the host (launch size, choice of what to mirror).
struct TailSmemLayout {
    uint32_t owners, hash_words, hash_keys, bid, queue0, total;
};
__host__ __device__ __forceinline__ TailSmemLayout tail_smem_layout(bool sprices, uint32_t own_mode, uint32_t n_cols, uint32_t cap) {
    TailSmemLayout l;
    uint32_t off = sprices ? ((n_cols * 8u + 15u) & ~15u) : 0u;
    l.owners = off;
    off += (own_mode == 1u) ? ((n_cols * 4u + 15u) & ~15u) : ((own_mode == 2u) ? ((n_cols * 2u + 15u) & ~15u) : 0u);
    l.hash_words = off;
    off += 2u * cap * 8u;
    l.hash_keys = off;
    off += 2u * cap * 4u;
    l.bid = off;
    off += cap * 8u;
    l.queue0 = off;
    off += 5u * cap * 4u;          // two queue buffers, object, previous owner, hash slot
    l.total = off;
    return l;
}

// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f64_order_key(double x) {
    unsigned long long u = (unsigned long long)__double_as_longlong(x);
    return u ^ ((u >> 63) ? ~0ull : (1ull << 63));
}
__host__ __device__ __forceinline__ double order_key_to_f64_host(unsigned long long k) {
    unsigned long long u = (k >> 63) ? (k ^ (1ull << 63)) : ~k;
    double d;
#ifdef __CUDA_ARCH__
    d = __longlong_as_double((long long)u);
#else
    __builtin_memcpy(&d, &u, 8);
#endif
    return d;
}

// Packed bid word: [63] = 0, then the top (63 - pbits) bits of the order key of the bid, then (pmask - person)
// so that on equal keys the LOWER person id wins.  Matches oracle/jacobi_model.c:jm_pack_bid bit for bit.
__device__ __forceinline__ unsigned long long pack_bid(double bid, uint32_t person, uint32_t pbits) {
    unsigned long long key = f64_order_key(bid);
    unsigned long long pmask = (1ull << pbits) - 1ull;
    return ((key >> (pbits + 1)) << pbits) | (pmask - (unsigned long long)person);
}

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xFFF0000000000000ll); }
__device__ __forceinline__ bool is_finite_f64(double x) {
    return (((unsigned long long)__double_as_longlong(x) >> 52) & 0x7FFull) != 0x7FFull;
}

// ---- loads ------------------------------------------------------------------------------------------------
// CSR arcs are streamed once per round: read-only path, do not allocate in L1.
__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ld_stream_d2(const double* p) {
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// Coherent, L1-allocating loads for state that the SAME CTA mutates between barriers (single-CTA engines):
// within one SM the L1 is kept coherent with that SM's own stores, so after __syncthreads() these see them.
// (Never used for words that are the target of L2 atomics.)
__device__ __forceinline__ double ld_ca_f64(const double* p) {
    double r;
    asm volatile("ld.global.ca.f64 %0, [%1];" : "=d"(r) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t ld_ca_u32(const uint32_t* p) {
    uint32_t r;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory");
    return r;
}

// Explicit shared-space accesses through 32-bit addresses.  On sm_100 the address of a shared variable embeds the
// CTA's rank in its cluster, which the compiler re-derives (S2UR SR_CgaCtaId + uniform arithmetic, tens of cycles on
// the critical path) wherever register pressure makes it drop the base; the latency-bound engines therefore compute
// their bases once (volatile, so they cannot be rematerialised) and address relative to them.
__device__ __forceinline__ uint32_t smem_base_u32(const void* p) {
    unsigned long long a;
    asm volatile("cvta.to.shared.u64 %0, %1;" : "=l"(a) : "l"(p));
    return (uint32_t)a;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {
    double r;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t a) {
    unsigned long long r;
    asm volatile("ld.shared.u64 %0, [%1];" : "=l"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t r;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ uint4 lds_u128(uint32_t a) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
    uint32_t r;
    asm volatile("{\n .reg .u16 t;\n ld.shared.u16 t, [%1];\n cvt.u32.u16 %0, t;\n}" : "=r"(r) : "r"(a) : "memory");
    return r;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) { asm volatile("st.shared.f64 [%0], %1;" :: "r"(a), "d"(v) : "memory"); }
__device__ __forceinline__ void sts_u64(uint32_t a, unsigned long long v) { asm volatile("st.shared.u64 [%0], %1;" :: "r"(a), "l"(v) : "memory"); }
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_u16(uint32_t a, uint32_t v) {
    asm volatile("{\n .reg .u16 t;\n cvt.u16.u32 t, %1;\n st.shared.u16 [%0], t;\n}" :: "r"(a), "r"(v) : "memory");
}
// First caller wins (small rounds: which of the equally informed last bidders publishes the totals).
__device__ __forceinline__ bool s_fin_claim(uint32_t* flag) { return atomicExch(flag, 1u) == 0u; }
// Barrier among the first-class participants of a small round only (warps whose slot is empty sleep at barrier 0).
__device__ __forceinline__ void named_bar_sync(uint32_t nthreads) {
    asm volatile("bar.sync 1, %0;" :: "r"(nthreads) : "memory");
}

// Where the tail engine reads the current owner of an object from: the global o2p array (coherent L1 loads) or a
// shared-memory mirror kept for the lifetime of the launch (u32, or u16 with 0xFFFF = none when N <= 65535).
struct OwnerView {
    const uint32_t* g;
    uint32_t* s32;
    uint16_t* s16;
    __device__ __forceinline__ bool mirrored() const { return s32 != nullptr || s16 != nullptr; }
    __device__ __forceinline__ uint32_t load(uint32_t j) const {
        if (s32) return s32[j];
        if (s16) { const uint32_t v = s16[j]; return v == 0xFFFFu ? SLA_DEV_NONE : v; }
        return ld_ca_u32(g + j);
    }
    __device__ __forceinline__ void store_mirror(uint32_t j, uint32_t person) const {
        if (s32) s32[j] = person;
        else if (s16) s16[j] = (uint16_t)person;
    }
};

// 256-bit streaming loads (sm_100): one LDG.256 per 8 column indices / 4 values, evict-first in L2 so that the
// once-per-round CSR stream does not push the object state (prices, owners, bid words) out of the L2.
__device__ __forceinline__ void ld_stream_u8(const uint32_t* p, uint32_t (&r)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "l"(p));
}
__device__ __forceinline__ void ld_stream_d4(const double* p, double* r) {
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r[0]), "=d"(r[1]), "=d"(r[2]), "=d"(r[3]) : "l"(p));
}

__device__ __forceinline__ uint4 ld_stream_u16x8(const uint16_t* p) {
    uint4 r;
    // (the L2::evict_first qualifier exists for the 256-bit forms only)
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// Exact u16 -> f64: 2^52 + x has x in its low mantissa bits, and the subtraction is exact.
__device__ __forceinline__ double u16_to_f64(uint32_t x) {
    return __hiloint2double(0x43300000, (int)x) - 4503599627370496.0;
}

enum PriceMode : int {
    PRICE_ZERO = 0,   // all prices are exactly 0.0: no gather
    PRICE_LDG = 1,    // prices immutable during this kernel: read-only path
    PRICE_CG = 2,     // L2-coherent loads
    PRICE_CA = 3,     // prices mutated by this very CTA (tail / batch engines): coherent L1-cached loads
    PRICE_SMEM = 4,   // `prices` points at a shared-memory copy kept by a single-CTA engine
    PRICE_MESH = 5,   // `prices` points at a MeshView: the price lives in the owner rank's HBM (peer-mapped over NVLink)
    PRICE_MESH_CG = 6 // the same read past the L1: persistent kernels, where prices change between the rounds of one launch
};

// Mesh engine (sla_mesh.cuh): object state is owner-partitioned over up to kMeshMaxRanks GPUs.  Price and owner of an
// object share one 16-byte cell (a remote price gather and an owner update touch the same sector); the packed bid words
// of the current round live in a separate, owner-local u64 array -- 8 B per object, so that the election of a round
// (one atomicMax and one read per bid) works on a quarter of the footprint.
constexpr int kMeshMaxRanks = 8;
struct alignas(16) ObjCell {
    double price;
    uint32_t owner;            // global person id or SLA_DEV_NONE
    uint32_t pad;
};
static_assert(sizeof(ObjCell) == 16, "one object cell is one 128-bit word");
struct MeshView {
    const ObjCell* cells[kMeshMaxRanks];   // rank g's cells (peer-mapped), object j lives at cells[j >> shift][j & mask]
    uint32_t shift, mask;
};

template <int MODE>
__device__ __forceinline__ double ld_price(const double* prices, uint32_t j) {
    if (MODE == PRICE_ZERO) return 0.0;
    if (MODE == PRICE_LDG) return __ldg(prices + j);   // L1-allocating on purpose: no_allocate gathers measured 1.6x slower
    if (MODE == PRICE_CA) return ld_ca_f64(prices + j);
    if (MODE == PRICE_SMEM) return prices[j];
    if (MODE == PRICE_MESH) {
        // frozen for the duration of the bid kernel, so the read-only path (L1) may cache it; peer addresses bypass the
        // local L2 and are served by the owner's
        const MeshView* mv = reinterpret_cast<const MeshView*>(prices);
        return __ldg(&(mv->cells[j >> mv->shift] + (j & mv->mask))->price);
    }
    if (MODE == PRICE_MESH_CG) {
        const MeshView* mv = reinterpret_cast<const MeshView*>(prices);
        return __ldcg(&(mv->cells[j >> mv->shift] + (j & mv->mask))->price);
    }
    return __ldcg(prices + j);
}

// ---- the choice rule ---------------------------------------------------------------------------------------
struct Choice {
    double best;     // max profit
    double second;   // second max profit (multiset sense)
    double value;    // edge value of the best arc
    uint32_t pos;    // global arc index of the best arc (lowest index wins ties) or SLA_DEV_NONE
    uint32_t col;    // its column
    uint32_t aux;    // single-CTA engines: current owner of the chosen column, filled in after the reduction
};

__device__ __forceinline__ void choice_init(Choice& c) {
    c.best = neg_inf(); c.second = neg_inf(); c.value = neg_inf(); c.pos = SLA_DEV_NONE; c.col = 0u; c.aux = SLA_DEV_NONE;
}

// Sequential update with one arc: exactly the reference's if / else-if (strict '>'), written with selects so that
// it compiles to straight-line code (the branchy form costs a divergence region per arc).
__device__ __forceinline__ void choice_update(Choice& c, double profit, double value, uint32_t pos, uint32_t col,
                                              uint32_t aux = SLA_DEV_NONE) {
    const bool gt = profit > c.best;
    const bool gs = profit > c.second;
    c.second = gt ? c.best : (gs ? profit : c.second);
    c.best = gt ? profit : c.best;
    c.value = gt ? value : c.value;
    c.pos = gt ? pos : c.pos;
    c.col = gt ? col : c.col;
    c.aux = gt ? aux : c.aux;
}

// Merge of two partial scans over disjoint arc sets; equals the sequential scan over their union in
// position order: max profit, lowest position among the maxima, second = second largest of the multiset.
__device__ __forceinline__ void choice_merge(Choice& c, double ob, double os, double ov, uint32_t opos, uint32_t ocol,
                                             uint32_t oaux) {
    const bool take = (ob > c.best) || (ob == c.best && opos < c.pos);
    const double loser_best = take ? c.best : ob;
    double s = (c.second > os) ? c.second : os;
    s = (loser_best > s) ? loser_best : s;
    c.best = take ? ob : c.best;
    c.value = take ? ov : c.value;
    c.pos = take ? opos : c.pos;
    c.col = take ? ocol : c.col;
    c.aux = take ? oaux : c.aux;
    c.second = s;
}

template <int LPR>
__device__ __forceinline__ void choice_group_reduce(Choice& c) {
#pragma unroll
    for (int m = LPR / 2; m >= 1; m >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, c.best, m);
        const double os = __shfl_xor_sync(0xffffffffu, c.second, m);
        const double ov = __shfl_xor_sync(0xffffffffu, c.value, m);
        const uint32_t op = __shfl_xor_sync(0xffffffffu, c.pos, m);
        const uint32_t oc = __shfl_xor_sync(0xffffffffu, c.col, m);
        const uint32_t oa = __shfl_xor_sync(0xffffffffu, c.aux, m);   // dead code where aux is never consumed
        choice_merge(c, ob, os, ov, op, oc, oa);
    }
}

// Scan of row [a, b) by a group of LPR lanes; each lane takes aligned chunks of 4 arcs (one 128-bit load of
// column indices, two 128-bit loads of values).  Arrays are padded so that the aligned chunk is always
// in bounds.  `flip` is 0 or 0x80000000 (sign normalisation of solver.rs:209-216 applied on the fly).
// STREAM: the CSR is read once per kernel (wide rounds) -> do not allocate in L1; otherwise (persistent single-CTA
// engines, where the same persons bid again and again) let the rows live in L1.
template <int LPR, int MODE, bool STREAM = true>
__device__ __forceinline__ void scan_row(Choice& c, const uint32_t* __restrict__ cols, const double* __restrict__ vals,
                                         const double* prices, uint32_t a, uint32_t b, uint32_t flip, int lane) {
    for (uint32_t base = (a & ~3u) + 4u * (uint32_t)lane; base < b; base += 4u * LPR) {
        uint4 cj;
        double2 v01, v23;
        if (STREAM) {
            cj = ld_stream_u4(cols + base);
            v01 = ld_stream_d2(vals + base);
            v23 = ld_stream_d2(vals + base + 2);
        } else {
            cj = __ldg(reinterpret_cast<const uint4*>(cols + base));
            v01 = __ldg(reinterpret_cast<const double2*>(vals + base));
            v23 = __ldg(reinterpret_cast<const double2*>(vals + base + 2));
        }
        const uint32_t jj[4] = {cj.x, cj.y, cj.z, cj.w};
        double vv[4] = {v01.x, v01.y, v23.x, v23.y};
        double pr[4];
        bool ok[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t g = base + t;
            ok[t] = (g >= a) && (g < b);
            pr[t] = ok[t] ? ld_price<MODE>(prices, jj[t]) : 0.0;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const double v = __hiloint2double(__double2hiint(vv[t]) ^ (int)flip, __double2loint(vv[t]));
            // arcs outside [a, b) get profit -inf, which never passes the strict '>' tests
            choice_update(c, ok[t] ? (v - pr[t]) : neg_inf(), v, base + t, jj[t]);
        }
    }
}

// Regular CSR fast path: 8 consecutive arcs starting at the 8-aligned arc index g, no masking.
template <int MODE>
__device__ __forceinline__ void scan8(Choice& c, const uint32_t* __restrict__ cols, const double* __restrict__ vals,
                                      const double* prices, uint32_t g, uint32_t flip) {
    uint32_t cj[8];
    double vv[8], pr[8];
    ld_stream_u8(cols + g, cj);
    ld_stream_d4(vals + g, vv);
    ld_stream_d4(vals + g + 4, vv + 4);
#pragma unroll
    for (int t = 0; t < 8; ++t) pr[t] = ld_price<MODE>(prices, cj[t]);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const double v = __hiloint2double(__double2hiint(vv[t]) ^ (int)flip, __double2loint(vv[t]));
        choice_update(c, (MODE == PRICE_ZERO) ? v : (v - pr[t]), v, g + t, cj[t]);
    }
}

// The same over the u16 mirror of the values (narrow upload, every value an integer in [0, 65535]): 6 instead of 12
// bytes per arc.  (double)u16 equals the widened f64 value bit for bit, so the choice is identical.
template <int MODE>
__device__ __forceinline__ void scan8_narrow(Choice& c, const uint32_t* __restrict__ cols, const uint16_t* __restrict__ vals16,
                                             const double* prices, uint32_t g, uint32_t flip) {
    uint32_t cj[8];
    double pr[8];
    ld_stream_u8(cols + g, cj);
    const uint4 w = ld_stream_u16x8(vals16 + g);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int t = 0; t < 8; ++t) pr[t] = ld_price<MODE>(prices, cj[t]);
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const double x = u16_to_f64((ww[t >> 1] >> ((t & 1) * 16)) & 0xFFFFu);
        const double v = __hiloint2double(__double2hiint(x) ^ (int)flip, __double2loint(x));
        choice_update(c, (MODE == PRICE_ZERO) ? v : (v - pr[t]), v, g + t, cj[t]);
    }
}

// First round (all prices exactly 0) over the u16 mirror, in integers: profit == +-value, so the choice rule is a
// top-2 over order-preserving keys.  key = value (maximising) or 65535 - value (values negated on the fly: larger key
// <=> larger -value); packed = key << 15 | (32767 - position in the row): the maximum is the best profit at the
// lowest position (the reference's strict '>' in position order, ksparse.rs:206 / symmetric.rs:367), the second
// largest packed word carries the second best profit in the multiset sense.  Rows of up to 32,768 arcs.
struct KeyChoice {
    int best, second;     // packed words, -1 = none yet
    uint32_t col;         // column of `best`
};
__device__ __forceinline__ void key_choice_init(KeyChoice& c) { c.best = -1; c.second = -1; c.col = 0u; }

__device__ __forceinline__ void scan8_keys(KeyChoice& c, const uint32_t* __restrict__ cols, const uint16_t* __restrict__ vals16,
                                           uint32_t g, uint32_t local, uint32_t keyflip) {
    uint32_t cj[8];
    ld_stream_u8(cols + g, cj);
    const uint4 w = ld_stream_u16x8(vals16 + g);
    const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
    const int base = 32767 - (int)local;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const uint32_t x = (t & 1) ? (ww[t >> 1] >> 16) : (ww[t >> 1] & 0xFFFFu);
        const int pk = (int)((x ^ keyflip) << 15) + (base - t);
        c.col = (pk > c.best) ? cj[t] : c.col;
        const int lo = min(c.best, pk);
        c.best = max(c.best, pk);
        c.second = max(c.second, lo);
    }
}

template <int LPR>
__device__ __forceinline__ void key_choice_group_reduce(KeyChoice& c) {
#pragma unroll
    for (int m = LPR / 2; m >= 1; m >>= 1) {
        const int ob = __shfl_xor_sync(0xffffffffu, c.best, m);
        const int os = __shfl_xor_sync(0xffffffffu, c.second, m);
        const uint32_t oc = __shfl_xor_sync(0xffffffffu, c.col, m);
        c.col = (ob > c.best) ? oc : c.col;
        const int lo = min(c.best, ob);
        c.best = max(c.best, ob);
        c.second = max(max(c.second, os), lo);
    }
}

// The f64 choice the packed words stand for (rows of at least two arcs: both words are set).
__device__ __forceinline__ void key_choice_to_f64(Choice& out, const KeyChoice& c, uint32_t row_begin, uint32_t keyflip, uint32_t flip) {
    const double xb = u16_to_f64(((uint32_t)c.best >> 15) ^ keyflip);
    const double xs = u16_to_f64(((uint32_t)c.second >> 15) ^ keyflip);
    out.value = __hiloint2double(__double2hiint(xb) ^ (int)flip, __double2loint(xb));
    out.best = out.value;
    out.second = __hiloint2double(__double2hiint(xs) ^ (int)flip, __double2loint(xs));
    out.pos = row_begin + (32767u - ((uint32_t)c.best & 32767u));
    out.col = c.col;
    out.aux = SLA_DEV_NONE;
}

// ---- warp-per-bidder scan for the small rounds of the single-CTA engines ---------------------------------------
// A whole warp scans one row (lane l takes arcs l, l+32, ...) and agrees on the choice with REDUX reductions on
// order-preserving integer keys of the profits instead of a shuffle tree of f64 compares: the dependent chain is
// five REDUX instructions whatever the row length.  Equivalent to the sequential choice rule: keys compare like the
// f64 profits (-0.0 canonicalised to +0.0); NaN and -inf profits never pass the reference's strict '>' against the
// initial -inf, so they get key 0 = "no arc"; ties go to the lowest arc position.
struct WarpChoice {
    double best, second, value, price;   // price = frozen price of the best column
    uint32_t col, pos, owner;            // pos == SLA_DEV_NONE: no usable arc; owner = current owner of `col`
};

enum OwnerMode : int { OWN_NONE = 0, OWN_GLOBAL = 1, OWN_SMEM = 2 };

// Order-preserving key of a profit; 0 = "no arc" for NaN and -inf (they never pass the reference's strict '>' against
// the initial -inf).  All-integer on purpose: the FP64 pipe costs ~35 cycles per dependent op on the latency-bound
// single-CTA engines.  -0.0 is canonicalised to +0.0 so that keys compare exactly like the f64 values.
__device__ __forceinline__ unsigned long long profit_key(double profit) {
    unsigned long long u = (unsigned long long)__double_as_longlong(profit);
    u = (u == 0x8000000000000000ull) ? 0ull : u;
    const unsigned long long k = u ^ ((u >> 63) ? ~0ull : (1ull << 63));
    // valid keys: key(-inf) < k <= key(+inf); negative NaNs fall below, positive NaNs above
    return (k > 0x000FFFFFFFFFFFFFull && k <= 0xFFF0000000000000ull) ? k : 0ull;
}
__device__ __forceinline__ double key_profit(unsigned long long key) {
    return key ? order_key_to_f64_host(key) : neg_inf();
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
    const uint32_t hi = __reduce_max_sync(0xffffffffu, (uint32_t)(k >> 32));
    const uint32_t lo = __reduce_max_sync(0xffffffffu, ((uint32_t)(k >> 32) == hi) ? (uint32_t)k : 0u);
    return ((unsigned long long)hi << 32) | lo;
}

// Per-lane running top-2 of the profit keys seen so far (lane-local part of the warp-per-bidder scan).
struct LaneTop2 {
    unsigned long long k1, k2;
    uint32_t pos1, col1;
    double val1, pr1;
};
__device__ __forceinline__ void lane_top2_init(LaneTop2& t) {
    t.k1 = 0ull; t.k2 = 0ull; t.pos1 = SLA_DEV_NONE; t.col1 = 0u; t.val1 = neg_inf(); t.pr1 = 0.0;
}
__device__ __forceinline__ void lane_top2_update(LaneTop2& t, unsigned long long key, uint32_t g, uint32_t col, double v,
                                                 double pr) {
    const bool gt = key > t.k1, gs = key > t.k2;
    t.k2 = gt ? t.k1 : (gs ? key : t.k2);
    t.k1 = gt ? key : t.k1;
    t.pos1 = gt ? g : t.pos1;
    t.col1 = gt ? col : t.col1;
    t.val1 = gt ? v : t.val1;
    t.pr1 = gt ? pr : t.pr1;
}

// Warp-wide agreement on the choice from the 32 lane-local top-2 records.
template <int OWN>
__device__ __forceinline__ WarpChoice warp_choice_finish(const LaneTop2& t, const uint32_t* owners) {
    // speculative: the current owner of this lane's best column, issued before the reductions so that its latency
    // hides behind them (only the owning lane's value is used)
    uint32_t own1 = SLA_DEV_NONE;
    if (OWN == OWN_GLOBAL) { if (t.pos1 != SLA_DEV_NONE) own1 = ld_ca_u32(owners + t.col1); }
    else if (OWN == OWN_SMEM) { if (t.pos1 != SLA_DEV_NONE) own1 = owners[t.col1]; }

    const unsigned long long kmax = warp_max_u64(t.k1);
    const bool is_max = (kmax != 0ull) && (t.k1 == kmax);
    const uint32_t posmin = __reduce_min_sync(0xffffffffu, is_max ? t.pos1 : SLA_DEV_NONE);
    const bool is_owner = is_max && (t.pos1 == posmin);
    const uint32_t owner_ballot = __ballot_sync(0xffffffffu, is_owner);
    const unsigned long long ksecond = warp_max_u64(is_owner ? t.k2 : t.k1);

    WarpChoice r;
    r.best = key_profit(kmax);
    r.second = key_profit(ksecond);
    r.pos = posmin;
    const int src = owner_ballot ? (__ffs((int)owner_ballot) - 1) : 0;
    r.value = __shfl_sync(0xffffffffu, t.val1, src);
    r.price = __shfl_sync(0xffffffffu, t.pr1, src);
    r.col = __shfl_sync(0xffffffffu, t.col1, src);
    r.owner = __shfl_sync(0xffffffffu, own1, src);
    if (!owner_ballot) { r.value = neg_inf(); r.price = 0.0; r.col = 0u; r.owner = SLA_DEV_NONE; r.pos = SLA_DEV_NONE; }
    return r;
}

template <int MODE, int OWN>
__device__ __forceinline__ WarpChoice warp_bid_scan(const uint32_t* __restrict__ cols, const double* __restrict__ vals,
                                                    const double* prices, const uint32_t* owners, uint32_t a, uint32_t b,
                                                    uint32_t flip, int lane32) {
    LaneTop2 t;
    lane_top2_init(t);
    for (uint32_t g = a + (uint32_t)lane32; g < b; g += 32u) {
        const uint32_t col = __ldg(cols + g);
        const double raw = __ldg(vals + g);
        const double v = __hiloint2double(__double2hiint(raw) ^ (int)flip, __double2loint(raw));
        const double pr = ld_price<MODE>(prices, col);
        lane_top2_update(t, profit_key((MODE == PRICE_ZERO) ? v : (v - pr)), g, col, v, pr);
    }
    return warp_choice_finish<OWN>(t, owners);
}

// ---- register-resident rows (slot-stable small rounds of the tail engine) --------------------------------------------
// A row of at most 32 * RPL arcs held by one warp: lane l keeps arcs a + l + 32 t (t < RPL) with the RAW uploaded
// values (the sign is applied where a value is used: a fetched row must not be touched before it is needed, or the
// fetch stops being asynchronous); slots past the end of the row are flagged in `len` terms by the user.
template <int RPL>
struct RowRegs {
    uint32_t c[RPL];
    double v[RPL];
    uint32_t a, len;
};

template <int RPL>
__device__ __forceinline__ void row_extents(RowRegs<RPL>& r, const uint32_t* __restrict__ row_ptr, uint32_t regK, uint32_t person) {
    if (regK) { r.a = person * regK; r.len = regK; }
    else { r.a = __ldg(row_ptr + person); r.len = __ldg(row_ptr + person + 1) - r.a; }
}

// `cols_lane` / `vals_lane` are the array bases already advanced by the lane index.
template <int RPL>
__device__ __forceinline__ void row_load(RowRegs<RPL>& r, const uint32_t* __restrict__ cols_lane, const double* __restrict__ vals_lane,
                                         int lane32) {
    const uint32_t* pc = cols_lane + r.a;
    const double* pv = vals_lane + r.a;
#pragma unroll
    for (int t = 0; t < RPL; ++t) {
        const uint32_t off = (uint32_t)lane32 + 32u * (uint32_t)t;
        r.c[t] = 0u;
        r.v[t] = 0.0;
        if (off < r.len && r.len <= 32u * RPL) {
            r.c[t] = __ldg(pc + 32 * t);
            r.v[t] = __ldg(pv + 32 * t);
        }
    }
}

template <int RPL, int MODE>
__device__ __forceinline__ LaneTop2 lane_scan_regs(const RowRegs<RPL>& r, const double* prices, uint32_t flip, int lane32) {
    LaneTop2 t;
    lane_top2_init(t);
    double pr[RPL];
#pragma unroll
    for (int u = 0; u < RPL; ++u) pr[u] = ld_price<MODE>(prices, r.c[u]);
#pragma unroll
    for (int u = 0; u < RPL; ++u) {
        const uint32_t off = (uint32_t)lane32 + 32u * (uint32_t)u;
        const double v = __hiloint2double(__double2hiint(r.v[u]) ^ (int)flip, __double2loint(r.v[u]));
        // arcs past the end of the row get key 0 = "no arc"
        lane_top2_update(t, (off < r.len) ? profit_key(v - pr[u]) : 0ull, r.a + off, r.c[u], v, pr[u]);
    }
    return t;
}

// Same scan with the price mirror addressed explicitly in shared space (`price_saddr` = shared address of price 0).
template <int RPL>
__device__ __forceinline__ LaneTop2 lane_scan_regs_smem(const RowRegs<RPL>& r, uint32_t price_saddr, uint32_t flip, int lane32) {
    LaneTop2 t;
    lane_top2_init(t);
    double pr[RPL];
#pragma unroll
    for (int u = 0; u < RPL; ++u) pr[u] = lds_f64(price_saddr + 8u * r.c[u]);
#pragma unroll
    for (int u = 0; u < RPL; ++u) {
        const uint32_t off = (uint32_t)lane32 + 32u * (uint32_t)u;
        const double v = __hiloint2double(__double2hiint(r.v[u]) ^ (int)flip, __double2loint(r.v[u]));
        lane_top2_update(t, (off < r.len) ? profit_key(v - pr[u]) : 0ull, r.a + off, r.c[u], v, pr[u]);
    }
    return t;
}

// Outcome of one person's bid: object (SLA_DEV_NONE = dropped by the Khosla threshold) and exact bid.
struct Bid {
    uint32_t obj;
    double bid;
    bool dropped;
};

// Bid computation from a finished choice (lane 0 of the group).
//   Khosla : ksparse.rs:218-227 (threshold test on the current price, `+= eps` when no finite second profit)
//   Forward: symmetric.rs:378
template <int MODE>
__device__ __forceinline__ Bid make_bid(const Choice& c, uint32_t algo, double eps, double threshold, const double* prices) {
    Bid r;
    r.obj = (c.pos == SLA_DEV_NONE) ? 0u : c.col;   // the reference starts from object 0 (ksparse.rs:196, symmetric.rs:355)
    r.dropped = false;
    if (algo == ALGO_KHOSLA) {
        const double pj = ld_price<MODE>(prices, r.obj);
        if (pj > threshold) { r.dropped = true; r.bid = 0.0; return r; }
        r.bid = is_finite_f64(c.second) ? (c.value - c.second + eps) : (pj + eps);
    } else {
        r.bid = c.value - c.second + eps;
    }
    return r;
}


// Bid from a WarpChoice (same expressions as make_bid): Khosla ksparse.rs:218-227, Forward symmetric.rs:378.
template <int MODE, int OWN>
__device__ __forceinline__ Bid make_bid_warp(const WarpChoice& c, uint32_t algo, double eps, double threshold,
                                             const double* prices, const uint32_t* owners, uint32_t* owner_out) {
    Bid r;
    r.dropped = false;
    double pj = c.price;
    uint32_t own = c.owner;
    if (c.pos == SLA_DEV_NONE) {   // row without a usable arc: the reference stays on object 0 (ksparse.rs:196, symmetric.rs:355)
        r.obj = 0u;
        pj = ld_price<MODE>(prices, 0u);
        own = (OWN == OWN_GLOBAL) ? ld_ca_u32(owners) : ((OWN == OWN_SMEM) ? owners[0] : SLA_DEV_NONE);
    } else {
        r.obj = c.col;
    }
    *owner_out = own;
    if (algo == ALGO_KHOSLA) {
        if (pj > threshold) { r.dropped = true; r.bid = 0.0; return r; }
        r.bid = is_finite_f64(c.second) ? (c.value - c.second + eps) : (pj + eps);
    } else {
        r.bid = c.value - c.second + eps;
    }
    return r;
}

// Bid from a WarpChoice whose owner has been settled by the caller (same expressions as make_bid): Khosla
// ksparse.rs:218-227, Forward symmetric.rs:378.  A row without a usable arc stays on object 0 (ksparse.rs:196,
// symmetric.rs:355).
template <int MODE>
__device__ __forceinline__ Bid make_bid_choice(const WarpChoice& c, uint32_t algo, double eps, double threshold,
                                               const double* prices) {
    Bid r;
    r.dropped = false;
    double pj = c.price;
    if (c.pos == SLA_DEV_NONE) {
        r.obj = 0u;
        pj = ld_price<MODE>(prices, 0u);
    } else {
        r.obj = c.col;
    }
    if (algo == ALGO_KHOSLA) {
        if (pj > threshold) { r.dropped = true; r.bid = 0.0; return r; }
        r.bid = is_finite_f64(c.second) ? (c.value - c.second + eps) : (pj + eps);
    } else {
        r.bid = c.value - c.second + eps;
    }
    return r;
}

}  // namespace sla
