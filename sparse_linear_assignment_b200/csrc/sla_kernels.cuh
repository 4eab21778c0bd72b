// sla_kernels.cuh -- the auction kernels: grid-wide bid scan / assign, the single-CTA tail engine, the
// Forward phase kernels (eps-CS check, reset), and the small utility kernels around them.
//
// Kernel order of one "super-round" (captured into a CUDA graph, DESIGN.md "Control flow"):
//     bid_wide -> assign_wide -> tail (control step A + tail rounds) -> [Forward only: ecs -> phase_apply]
// Every kernel decides from the device-resident DevState whether it has work, so the same sequence can be
// replayed without host involvement until DevState::done is set.
#pragma once
#include <cstddef>
#include <cooperative_groups.h>
#include "sla_common.cuh"
#include "synth.h"

namespace sla {

// =============================================================================================================
// Grid-wide bid scan  (reference hot loops src/ksparse.rs:199-227, src/symmetric.rs:343-384)
// One group of LPR lanes per bidder; coalesced 128-bit loads of the CSR row; shuffle reduction of
// (best, second best, best value, best position); one 64-bit atomicMax per bid for conflict resolution.
// =============================================================================================================
template <int LPR, int MODE>
__device__ __forceinline__ void bid_wide_body(const Params& p, const uint32_t qlen, const bool identity,
                                              const uint32_t* __restrict__ queue, const uint32_t algo, const double eps,
                                              const double threshold, const uint32_t pbits, const uint32_t sign_flip,
                                              const uint32_t person_base) {
    constexpr int GROUPS_PER_BLOCK = kWideThreads / LPR;
    const int lane = threadIdx.x % LPR;
    const uint32_t group = blockIdx.x * GROUPS_PER_BLOCK + threadIdx.x / LPR;
    const uint32_t ngroups = gridDim.x * GROUPS_PER_BLOCK;
    unsigned long long my_arcs = 0;
    uint32_t my_dropped = 0;

    const uint32_t warp_group0 = group - (uint32_t)((threadIdx.x & 31) / LPR);   // first group of this warp
    for (uint32_t base = 0; base < qlen; base += ngroups) {
        if (base + warp_group0 >= qlen) break;   // warp-uniform: the whole warp is past the end of the queue
        const uint32_t q = base + group;
        const bool valid = q < qlen;
        uint32_t i = 0, a = 0, b = 0;
        if (valid) {
            i = identity ? q : __ldg(queue + q);
            a = __ldg(p.row_ptr + i);
            b = __ldg(p.row_ptr + i + 1);
        }
        Choice c;
        choice_init(c);
        scan_row<LPR, MODE>(c, p.cols, p.vals, p.prices, a, b, sign_flip, lane);
        choice_group_reduce<LPR>(c);
        if (valid && lane == 0) {
            const Bid r = make_bid<MODE>(c, algo, eps, threshold, p.prices);
            my_arcs += (unsigned long long)(b - a);
            if (r.dropped) {
                p.slot_obj[q] = SLA_DEV_NONE;
                my_dropped += 1;
            } else {
                p.slot_obj[q] = r.obj;
                p.slot_bid[q] = r.bid;
                if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, i + person_base, pbits));   // NaN never bids
            }
        }
    }

    // per-block accumulation of the instrumentation counters
    __shared__ unsigned long long s_arcs;
    __shared__ uint32_t s_dropped;
    if (threadIdx.x == 0) { s_arcs = 0; s_dropped = 0; }
    __syncthreads();
    if (my_arcs) atomicAdd(&s_arcs, my_arcs);
    if (my_dropped) atomicAdd(&s_dropped, my_dropped);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_arcs) atomicAdd(&p.st->bid_arcs, s_arcs);
        if (s_dropped) atomicAdd(&p.st->dropped, s_dropped);
    }
}

template <int LPR>
__global__ void __launch_bounds__(kWideThreads) bid_wide_kernel(const Params p) {
    const HotState h = load_hot(p.st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (h.done || qlen <= h.tail_max) return;
    const bool identity = h.identity != 0;
    const bool zero = (h.zero_prices != 0) && (h.skip_zero != 0);
    const uint32_t algo = h.algo, pbits = h.pbits;
    const double eps = h.eps, thr = h.threshold;
    const uint32_t sf = h.sign_flip;
    if (zero) bid_wide_body<LPR, PRICE_ZERO>(p, qlen, identity, cur ? p.queue[1] : p.queue[0], algo, eps, thr, pbits, sf, h.person_base);
    else      bid_wide_body<LPR, PRICE_LDG>(p, qlen, identity, cur ? p.queue[1] : p.queue[0], algo, eps, thr, pbits, sf, h.person_base);
}

// One row of a uniform-degree CSR (K <= 8 * LPR8: one 8-arc chunk per lane) through the bound-pruned gather: see
// bid_regular_body.  All LPR8 lanes of the group call it (also for an invalid slot); the finished choice is in every lane.
template <int LPR8, int MODE, bool NARROW>
__device__ __forceinline__ void scan_row_pruned(Choice& c, const Params& p, const double* prices, const uint32_t i,
                                                const bool valid, const uint32_t K, const uint32_t sign_flip, const int lane) {
    const uint32_t keyflip = sign_flip ? 0xFFFFu : 0u;
    const uint32_t off = 8u * (uint32_t)lane;
    const bool has = valid && off < K;
    uint32_t cj[8];
    double vv[8];
    int key[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { cj[t] = 0u; vv[t] = neg_inf(); key[t] = (int)0x80000000; }
    uint32_t g = 0;
    if (has) {
        g = i * K + off;
        ld_stream_u8(p.cols + g, cj);
        if (NARROW) {
            const uint4 w = ld_stream_u16x8(p.vals16 + g);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const uint32_t x = (t & 1) ? (ww[t >> 1] >> 16) : (ww[t >> 1] & 0xFFFFu);
                const double d = u16_to_f64(x);
                vv[t] = __hiloint2double(__double2hiint(d) ^ (int)sign_flip, __double2loint(d));
                key[t] = (int)(x ^ keyflip);                 // exact order of the effective values
            }
        } else {
            double raw[8];
            ld_stream_d4(p.vals + g, raw);
            ld_stream_d4(p.vals + g + 4, raw + 4);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int hi = __double2hiint(raw[t]) ^ (int)sign_flip;
                vv[t] = __hiloint2double(hi, __double2loint(raw[t]));
                key[t] = hi ^ ((hi >> 31) & 0x7FFFFFFF);     // order of the high words: any two probes give a valid bound
            }
        }
    }
    // the lane's two probes: (approximately) its two most valuable arcs -- a top-2 over `key with the position in its low
    // three bits` (integer min / max, three instructions per arc; which two arcs are probed does not matter for the result)
    int b1 = (int)0x80000000, b2 = (int)0x80000000;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const int pk = NARROW ? (int)(((uint32_t)key[t] << 3) | (uint32_t)t) : ((key[t] & ~7) | t);
        const int lo = min(b1, pk);
        b1 = max(b1, pk);
        b2 = max(b2, lo);
    }
    const uint32_t i1 = (uint32_t)b1 & 7u, i2 = (uint32_t)b2 & 7u;      // distinct: every pk carries its own position
    double pr[8];
    double hiP = neg_inf(), loP = neg_inf();
    {
        double mx = neg_inf(), mn = __longlong_as_double(0x7FF0000000000000ll);
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const bool probe = has && (((uint32_t)t == i1) || ((uint32_t)t == i2));
            pr[t] = 0.0;
            if (probe) pr[t] = ld_price<MODE>(prices, cj[t]);
        }
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            const bool probe = ((uint32_t)t == i1) || ((uint32_t)t == i2);
            const double pf = vv[t] - pr[t];
            mx = (probe && pf > mx) ? pf : mx;
            mn = (probe && pf < mn) ? pf : mn;
        }
        if (has) { hiP = mx; loP = mn; }
    }
#pragma unroll
    for (int m = LPR8 / 2; m >= 1; m >>= 1) {
        const double oh = __shfl_xor_sync(0xffffffffu, hiP, m);
        const double ol = __shfl_xor_sync(0xffffffffu, loP, m);
        const double mn = hiP > oh ? oh : hiP;          // the smaller of the two maxima
        double l2 = loP > ol ? loP : ol;
        l2 = mn > l2 ? mn : l2;
        hiP = hiP > oh ? hiP : oh;
        loP = l2;
    }
    const double bound = loP;
    bool use[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const bool probe = ((uint32_t)t == i1) || ((uint32_t)t == i2);
        const bool more = has && !probe && vv[t] >= bound;
        use[t] = has && (probe || more);
        if (more) pr[t] = ld_price<MODE>(prices, cj[t]);
    }
    choice_init(c);
#pragma unroll
    for (int t = 0; t < 8; ++t)
        choice_update(c, use[t] ? (vv[t] - pr[t]) : neg_inf(), vv[t], g + t, cj[t]);
    choice_group_reduce<LPR8>(c);
}

#ifndef SLA_PRUNE_MIN_QUEUE
#define SLA_PRUNE_MIN_QUEUE 32768
#endif
constexpr uint32_t kPruneMinQueue = SLA_PRUNE_MIN_QUEUE;   // bidders from which the gathering scan prunes by value bound

// Regular-CSR variant (every row has exactly K arcs, K % 8 == 0: all of BASELINE.json's configs): no row-extent
// loads (a = i * K), no masking, 8 arcs per lane per step through 256-bit loads.  LPR8 lanes share one row.
template <int LPR8, int MODE, bool NARROW>
__device__ __forceinline__ void bid_regular_body(const Params& p, const uint32_t qlen, const bool identity,
                                                 const uint32_t* __restrict__ queue, const uint32_t algo, const double eps,
                                                 const double threshold, const uint32_t pbits, const uint32_t sign_flip,
                                                 const uint32_t K, const uint32_t person_base, const bool prune) {
    constexpr int GROUPS_PER_BLOCK = kWideThreads / LPR8;
    const int lane = threadIdx.x % LPR8;
    const uint32_t group = blockIdx.x * GROUPS_PER_BLOCK + threadIdx.x / LPR8;
    const uint32_t ngroups = gridDim.x * GROUPS_PER_BLOCK;
    uint32_t my_dropped = 0;

    const uint32_t warp_group0 = group - (uint32_t)((threadIdx.x & 31) / LPR8);   // first group of this warp
#ifndef SLA_REG_UNROLL
#define SLA_REG_UNROLL 2
#endif
#ifndef SLA_REG_ADJ
#define SLA_REG_ADJ 0
#endif
    // rows per group and pass: all their loads are in flight before the first reduction (streaming variant only)
#ifndef SLA_REG_UNROLL_LDG
#define SLA_REG_UNROLL_LDG 1
#endif
    constexpr int U = (MODE == PRICE_ZERO) ? SLA_REG_UNROLL : SLA_REG_UNROLL_LDG;
    constexpr bool ADJ = SLA_REG_ADJ != 0;   // the U rows of a group are neighbours in the queue
    if constexpr (NARROW && MODE == PRICE_ZERO) {
        // u16 values and all prices zero: the choice in integer keys (scan8_keys), converted to the f64 choice at the end
        // (the host takes this kernel for K <= 32,768 only: narrow_scan_ptr)
        const uint32_t keyflip = sign_flip ? 0xFFFFu : 0u;
#ifndef SLA_KEY_UNROLL
#define SLA_KEY_UNROLL 2
#endif
        constexpr int UK = SLA_KEY_UNROLL;           // rows per group and pass (48 B per lane and row in flight)
        for (uint32_t base = 0; base < qlen; base += ngroups * UK) {
            if (base + warp_group0 >= qlen) break;   // warp-uniform: the whole warp is past the end
            KeyChoice kc[UK];
            uint32_t i[UK];
            bool valid[UK];
#pragma unroll
            for (int u = 0; u < UK; ++u) {
                const uint32_t q = base + (uint32_t)u * ngroups + group;
                valid[u] = q < qlen;
                key_choice_init(kc[u]);
                i[u] = 0;
                if (valid[u]) {
                    i[u] = identity ? q : __ldg(queue + q);
                    const uint32_t a = i[u] * K;
                    for (uint32_t off = 8u * (uint32_t)lane; off < K; off += 8u * LPR8)
                        scan8_keys(kc[u], p.cols, p.vals16, a + off, off, keyflip);
                }
            }
#pragma unroll
            for (int u = 0; u < UK; ++u) {
                if (u && base + (uint32_t)u * ngroups + warp_group0 >= qlen) break;   // warp-uniform
                const uint32_t q = base + (uint32_t)u * ngroups + group;
                key_choice_group_reduce<LPR8>(kc[u]);
                if (valid[u] && lane == 0) {
                    Choice c;
                    key_choice_to_f64(c, kc[u], i[u] * K, keyflip, sign_flip);
                    const Bid r = make_bid<MODE>(c, algo, eps, threshold, p.prices);
                    if (r.dropped) {
                        p.slot_obj[q] = SLA_DEV_NONE;
                        my_dropped += 1;
                    } else {
                        p.slot_obj[q] = r.obj;
                        p.slot_bid[q] = r.bid;
                        if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, i[u] + person_base, pbits));
                    }
                }
            }
        }
    } else if (MODE == PRICE_LDG && prune && K <= 8u * LPR8) {
        // ---- bound-pruned gather ------------------------------------------------------------------------------------
        // Prices never go below zero (they start at 0 and every winning bid is >= old price + eps, eps >= 0: DevState::
        // prune_ok), so profit = value - price <= value.  Each lane first gathers the prices of its two most valuable
        // arcs only; the second largest of the 2 x LPR8 profits found that way, L, is a lower bound on the row's true
        // second-best profit.  An arc with value < L has profit < L <= second <= best: it can change neither the best
        // arc, nor the second-best profit, nor a tie (ties need equality) -- its price is not fetched.  Every other arc
        // is gathered and goes through the unchanged choice rule in position order, so the choice is the reference's
        // bit for bit (ksparse.rs:199-214, symmetric.rs:361-376).  What changes is the number of scattered 8-byte
        // gathers per row: 2 per lane plus the few arcs that can still matter, instead of all K -- the gathers (one L1
        // wavefront each), not the CSR stream, bound the unpruned scan.
        for (uint32_t base = 0; base < qlen; base += ngroups) {
            if (base + warp_group0 >= qlen) break;   // warp-uniform: the whole warp is past the end
            const uint32_t q = base + group;
            const bool valid = q < qlen;
            uint32_t i = 0;
            if (valid) i = identity ? q : __ldg(queue + q);
            Choice c;
            scan_row_pruned<LPR8, MODE, NARROW>(c, p, p.prices, i, valid, K, sign_flip, lane);
            if (valid && lane == 0) {
                const Bid r = make_bid<MODE>(c, algo, eps, threshold, p.prices);
                if (r.dropped) {
                    p.slot_obj[q] = SLA_DEV_NONE;
                    my_dropped += 1;
                } else {
                    p.slot_obj[q] = r.obj;
                    p.slot_bid[q] = r.bid;
                    if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, i + person_base, pbits));
                }
            }
        }
    } else {
    for (uint32_t base = 0; base < qlen; base += ngroups * U) {
        if (base + (ADJ ? warp_group0 * U : warp_group0) >= qlen) break;   // warp-uniform: the whole warp is past the end
        Choice c[U];
        uint32_t i[U];
        bool valid[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t q = ADJ ? base + group * U + (uint32_t)u : base + (uint32_t)u * ngroups + group;
            valid[u] = q < qlen;
            choice_init(c[u]);
            i[u] = 0;
            if (valid[u]) {
                i[u] = identity ? q : __ldg(queue + q);
                const uint32_t a = i[u] * K;
                for (uint32_t off = 8u * (uint32_t)lane; off < K; off += 8u * LPR8) {
                    if (NARROW) scan8_narrow<MODE>(c[u], p.cols, p.vals16, p.prices, a + off, sign_flip);
                    else scan8<MODE>(c[u], p.cols, p.vals, p.prices, a + off, sign_flip);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ADJ && u && base + (uint32_t)u * ngroups + warp_group0 >= qlen) break;   // warp-uniform
            const uint32_t q = ADJ ? base + group * U + (uint32_t)u : base + (uint32_t)u * ngroups + group;
            choice_group_reduce<LPR8>(c[u]);
            if (valid[u] && lane == 0) {
                const Bid r = make_bid<MODE>(c[u], algo, eps, threshold, p.prices);
                if (r.dropped) {
                    p.slot_obj[q] = SLA_DEV_NONE;
                    my_dropped += 1;
                } else {
                    p.slot_obj[q] = r.obj;
                    p.slot_bid[q] = r.bid;
#if defined(SLA_EXP_RED) && SLA_EXP_RED == 1
                    // experiment: no conflict-resolution atomics at all (results meaningless)
#elif defined(SLA_EXP_RED) && SLA_EXP_RED == 2
                    if (r.bid == r.bid) atomicMax(p.best + (r.obj & 0xFFFFu), pack_bid(r.bid, i[u] + person_base, pbits));
#else
                    if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, i[u] + person_base, pbits));
#endif
                }
            }
        }
    }
    }
    if (my_dropped) atomicAdd(&p.st->dropped, my_dropped);
    // arcs of a regular round = bidders * K: accounted in control step A
}

// Values that crossed PCIe as u16 / f32 (upload_large): restore the f64 array, 8 values per thread and pass.
template <class T>
__global__ void __launch_bounds__(kWideThreads) widen_values_kernel(const T* __restrict__ src, double* __restrict__ dst,
                                                                    const size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 8u;
    for (size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8u; g < n; g += stride) {
        if (g + 8u <= n) {
            T t[8];
            if (sizeof(T) == 2) *reinterpret_cast<uint4*>(t) = *reinterpret_cast<const uint4*>(src + g);
            else { *reinterpret_cast<uint4*>(t) = *reinterpret_cast<const uint4*>(src + g);
                   *reinterpret_cast<uint4*>(t + 4) = *reinterpret_cast<const uint4*>(src + g + 4); }
#pragma unroll
            for (int u = 0; u < 8; u += 2)
                *reinterpret_cast<double2*>(dst + g + u) = make_double2((double)t[u], (double)t[u + 1]);
        } else {
            for (size_t u = g; u < n; ++u) dst[u] = (double)src[u];
        }
    }
}

// u16 uploads: ONE launch per run of staged pieces, right behind the run's copy on the same stream, so that nothing but
// the last run's few microseconds is left to do when the last piece has landed: restores values [lo, hi) of the f64 array
// from their u16 image, folds their range into the statistics csr_stats_kernel would compute (for non-negative integers
// the order of the f64 keys is the order of the u16 images) and checks the column indices of the same arcs (the whole
// column array crossed PCIe in front of the first piece).  `lo` is a multiple of the piece size (64 Ki values).
__global__ void __launch_bounds__(kWideThreads) widen_u16_stats_kernel(const uint16_t* __restrict__ src, double* __restrict__ dst,
                                                                       const uint32_t* __restrict__ cols, const size_t lo,
                                                                       const size_t hi, const uint32_t n_cols, DevCsrStats* out) {
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    uint32_t mn = 0xFFFFu, mx = 0u;
    unsigned long long bad_c = 0ull;
    bool any = false;
    const size_t octs = (hi - lo) / 8u;
    for (size_t t = tid; t < octs; t += stride) {
        const size_t g = lo + 8u * t;
        const uint4 w = __ldg(reinterpret_cast<const uint4*>(src + g));
        const uint4 c0 = __ldg(reinterpret_cast<const uint4*>(cols + g)), c1 = __ldg(reinterpret_cast<const uint4*>(cols + g + 4));
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t a = ww[u] & 0xFFFFu, b = ww[u] >> 16;
            mn = min(mn, min(a, b));
            mx = max(mx, max(a, b));
            *reinterpret_cast<double2*>(dst + g + 2 * u) = make_double2((double)a, (double)b);
        }
        bad_c += (c0.x >= n_cols ? 1u : 0u) + (c0.y >= n_cols ? 1u : 0u) + (c0.z >= n_cols ? 1u : 0u) + (c0.w >= n_cols ? 1u : 0u) +
                 (c1.x >= n_cols ? 1u : 0u) + (c1.y >= n_cols ? 1u : 0u) + (c1.z >= n_cols ? 1u : 0u) + (c1.w >= n_cols ? 1u : 0u);
        any = true;
    }
    for (size_t g = lo + octs * 8u + tid; g < hi; g += stride) {
        const uint32_t a = src[g];
        mn = min(mn, a);
        mx = max(mx, a);
        dst[g] = (double)a;
        bad_c += (cols[g] >= n_cols) ? 1u : 0u;
        any = true;
    }
    const uint32_t have = __ballot_sync(0xffffffffu, any);
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) bad_c += __shfl_xor_sync(0xffffffffu, bad_c, m);
    if ((threadIdx.x & 31) == 0 && have) {
        atomicMin(&out->min_key, f64_order_key((double)mn));
        atomicMax(&out->max_key, f64_order_key((double)mx));
        if (bad_c) atomicAdd(&out->bad_cols, bad_c);
    }
}

// The row half of csr_stats_kernel (extents monotone and within nnz, uniform degree), for the same upload path.
__global__ void __launch_bounds__(kWideThreads) csr_row_stats_kernel(const uint32_t* __restrict__ row_ptr, const uint32_t n_rows,
                                                                     const unsigned long long nnz, DevCsrStats* out) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad_r = 0, irr = 0;
    const uint32_t k0 = row_ptr[1] - row_ptr[0];
    for (unsigned long long i = tid; i < n_rows; i += stride) {
        const uint32_t a = row_ptr[i], b = row_ptr[i + 1];
        bad_r += (b < a || (unsigned long long)b > nnz) ? 1u : 0u;
        irr += (b - a != k0) ? 1u : 0u;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        bad_r += __shfl_xor_sync(0xffffffffu, bad_r, m);
        irr += __shfl_xor_sync(0xffffffffu, irr, m);
    }
    if ((threadIdx.x & 31) == 0) {
        if (bad_r) atomicAdd(&out->bad_rows, bad_r);
        if (irr) atomicAdd(&out->irregular_rows, irr);
    }
}

// Empty kernel: the "profile" mode launches it in front of an event so that the event is recorded by the compute
// front-end right before the kernel it times (and not behind the copy engine's upload of the control block).
__global__ void profile_fence_kernel() {}

// MODE is a launch-time decision of the host: PRICE_ZERO only for the very first round of a solve (prices are
// exactly 0 after init_solve, solver.rs:218-219, and the option zero_price_skip is on), PRICE_LDG otherwise.
// NARROW: the values are read from their u16 mirror (Params::vals16), which a narrow upload left in HBM.
template <int LPR8, int MODE, bool NARROW = false>
#ifndef SLA_REG_MINB_ZERO
#define SLA_REG_MINB_ZERO 3
#endif
// CTAs per SM of the integer-key first round (46 registers at 5, spills from 6 up): cfg3 scan 29.4 us at 3, 26.7 at 4
#ifndef SLA_KEY_OCC
#define SLA_KEY_OCC 5
#endif
__global__ void __launch_bounds__(kWideThreads, (MODE == PRICE_ZERO) ? (NARROW ? SLA_KEY_OCC : SLA_REG_MINB_ZERO) : 4) bid_regular_kernel(const Params p) {
    const HotState h = load_hot(p.st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (h.done || qlen <= h.tail_max) return;
    const bool identity = h.identity != 0;
    const uint32_t algo = h.algo, pbits = h.pbits, sf = h.sign_flip, K = h.regular_k;
    const double eps = h.eps, thr = h.threshold;
    const uint32_t* queue = cur ? p.queue[1] : p.queue[0];
    // bound-pruned gather (bid_regular_body): two dependent gather phases instead of one, so only where the scan is
    // throughput-bound, not for the short latency-bound rounds
    const bool prune = (MODE == PRICE_LDG) && qlen >= kPruneMinQueue && p.st->prune_ok != 0u;
    bid_regular_body<LPR8, MODE, NARROW>(p, qlen, identity, queue, algo, eps, thr, pbits, sf, K, h.person_base, prune);
}

// =============================================================================================================
// First-round bid scan as a TMA pipeline (sm_100a): regular CSR, identity queue, all prices exactly zero.
// The rows of consecutive persons are one contiguous byte range of `cols` and of `vals`, so a tile of
// kStreamThreads / LPR8 rows is fetched with two bulk copies (cp.async.bulk global -> shared, completion counted on an
// mbarrier) by a producer warp that runs kStreamStages tiles ahead of the eight consumer warps.  The consumers apply
// exactly the choice rule of scan8 / choice_group_reduce / make_bid to the staged tile, so the results (slots, bid
// words) are bit-identical to bid_regular_kernel<LPR8, PRICE_ZERO>; what changes is who waits for HBM: the copy
// engine keeps kStreamStages x 24 KB per CTA in flight whatever the consumers are doing.
// =============================================================================================================
constexpr int kStreamThreads = 256;                 // consumer threads (8 warps); warp 8 is the producer
constexpr int kStreamStages = 4;
constexpr uint32_t kStreamTileArcs = kStreamThreads * 8;                 // 2048 arcs: 8 KB of columns + 16 KB of values
constexpr uint32_t kStreamStageBytes = kStreamTileArcs * 12u;
constexpr uint32_t kStreamSmemBytes = kStreamStages * kStreamStageBytes + 2u * kStreamStages * 8u;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy, evict-first in the L2 (the CSR streams past the object state exactly once per round)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}

template <int LPR8>
__global__ void __launch_bounds__(kStreamThreads + 32, 2) bid_stream_kernel(const Params p) {
    extern __shared__ __align__(128) unsigned char stream_smem[];
    const HotState h = load_hot(p.st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (h.done || qlen <= h.tail_max) return;
    if (!h.identity) __trap();   // host contract: launched for the first round of a solve only
    const uint32_t K = h.regular_k;
    constexpr uint32_t T = kStreamThreads / LPR8;          // rows per tile
    const uint32_t ntiles = (qlen + T - 1u) / T;
    const uint32_t bars = smem_u32(stream_smem + kStreamStages * kStreamStageBytes);   // full[s] at +8s, empty[s] after them

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kStreamStages; ++s) {
            mbar_init(bars + 8u * s, 1u);                                   // the producer's expect_tx arrival
            mbar_init(bars + 8u * (kStreamStages + s), kStreamThreads / 32);   // one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (threadIdx.x >= kStreamThreads) {
        // ---- producer warp: one lane issues the copies -----------------------------------------------------
        if (threadIdx.x == kStreamThreads) {
            uint64_t policy;
            asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
            uint32_t it = 0;
            for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const uint32_t s = it % kStreamStages, use = it / kStreamStages;
                if (use) mbar_wait(bars + 8u * (kStreamStages + s), (use - 1u) & 1u);   // consumers released use-1
                const uint32_t row0 = tile * T;
                const uint32_t rows = (qlen - row0 < T) ? (qlen - row0) : T;
                const uint32_t arcs = rows * K;
                const size_t a0 = (size_t)row0 * K;
                const uint32_t full = bars + 8u * s;
                const uint32_t base = smem_u32(stream_smem + s * kStreamStageBytes);
                mbar_expect_tx(full, arcs * 12u);
                bulk_g2s(base, p.cols + a0, arcs * 4u, full, policy);
                bulk_g2s(base + kStreamTileArcs * 4u, p.vals + a0, arcs * 8u, full, policy);
            }
        }
        return;
    }

    // ---- consumers: LPR8 lanes per row, 8 arcs per lane and pass (the layout of bid_regular_body) -----------------
    const int lane = threadIdx.x % LPR8;
    const uint32_t lrow = threadIdx.x / LPR8;
    const uint32_t algo = h.algo, pbits = h.pbits, sf = h.sign_flip, person_base = h.person_base;
    const double eps = h.eps, thr = h.threshold;
    uint32_t my_dropped = 0, it = 0;
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
        const uint32_t s = it % kStreamStages, use = it / kStreamStages;
        const uint32_t q = tile * T + lrow;
        const bool valid = q < qlen;
        const unsigned char* stage = stream_smem + s * kStreamStageBytes;
        const uint32_t* scols = reinterpret_cast<const uint32_t*>(stage);
        const double* svals = reinterpret_cast<const double*>(stage + kStreamTileArcs * 4u);
        Choice c;
        choice_init(c);
        mbar_wait(bars + 8u * s, use & 1u);
        if (valid) {
            for (uint32_t off = 8u * (uint32_t)lane; off < K; off += 8u * LPR8) {
                const uint32_t la = lrow * K + off;            // arc index inside the tile
                const uint4 c0 = *reinterpret_cast<const uint4*>(scols + la);
                const uint4 c1 = *reinterpret_cast<const uint4*>(scols + la + 4);
                const uint32_t cj[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                double vv[8];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const double2 d = *reinterpret_cast<const double2*>(svals + la + 2 * t);
                    vv[2 * t] = d.x; vv[2 * t + 1] = d.y;
                }
                const uint32_t g = q * K + off;                // global arc index (tie-break position)
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const double v = __hiloint2double(__double2hiint(vv[t]) ^ (int)sf, __double2loint(vv[t]));
                    choice_update(c, v, v, g + t, cj[t]);
                }
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(bars + 8u * (kStreamStages + s));   // this warp is done with the stage
        choice_group_reduce<LPR8>(c);
        if (valid && lane == 0) {
            const Bid r = make_bid<PRICE_ZERO>(c, algo, eps, thr, p.prices);
            if (r.dropped) {
                p.slot_obj[q] = SLA_DEV_NONE;
                my_dropped += 1;
            } else {
                p.slot_obj[q] = r.obj;
                p.slot_bid[q] = r.bid;
                if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, q + person_base, pbits));
            }
        }
    }
    if (my_dropped) atomicAdd(&p.st->dropped, my_dropped);
}

// =============================================================================================================
// Control helpers (executed by exactly one thread)
// =============================================================================================================
// Queue ran empty.  Forward is finished without an eps-CS check when start_from_optimal_eps holds
// (symmetric.rs:279-288); otherwise the ecs kernel decides.  Khosla is finished (ksparse.rs:186 loop exit) unless
// its rounds run under the eps-schedule (square instances): then a phase in which nobody was dropped is followed by
// the next, smaller eps (the assignment is wiped, the prices are kept, the last phase runs at exactly the caller's
// eps), and a phase that dropped anybody -- the instance is not known to have a perfect matching, and only the plain
// rounds define what the reference's price threshold does then -- makes the solve start over without a schedule.
// ST is DevState (the tail engine's shared-memory copy of the block) or volatile DevState (the block in global memory,
// read and written past the L1 by the last block of a grid-wide kernel).
template <class ST>
__device__ __forceinline__ void finish_if_possible(ST* st) {
    if (st->algo == ALGO_KHOSLA) {
        if (!st->kscale) { st->done = 1; return; }
        if (st->dropped != 0) {
            st->kscale = 0;
            st->eps = st->target_eps;
            st->dropped = 0;
            st->action = ACTION_RESET_ALL;
        } else if (st->eps > st->target_eps) {
            const double e = st->eps * 0.15, te = st->target_eps;
            st->eps = (e < te) ? te : e;
            st->nreductions = st->nreductions + 1;
            st->action = ACTION_RESET;
        } else {
            st->done = 1;
            return;
        }
        st->qlen[st->cur] = st->n_rows;
        st->identity = 1;
    } else if (st->start_opt) {
        st->optimal = 1;
        st->done = 1;
    }
}

// Control step A: account for the wide round of this super-round (if one ran) and flip the queues.
// Returns the current queue length afterwards.
template <class ST>
__device__ __forceinline__ uint32_t control_after_wide(ST* st) {
    uint32_t cur = st->cur;
    uint32_t qlen = st->qlen[cur];
    st->action = ACTION_NONE;
    if (!st->done && qlen > st->tail_max) {
        st->rounds = st->rounds + 1;
        st->wide_rounds = st->wide_rounds + 1;
        st->bids = st->bids + qlen;
        if (st->regular_k) st->bid_arcs = st->bid_arcs + (unsigned long long)qlen * st->regular_k;
        st->qlen[cur] = 0;
        cur ^= 1u;
        st->cur = cur;
        qlen = st->qlen[cur];
        st->identity = 0;
        st->zero_prices = 0;
        if (st->algo == ALGO_FORWARD) {
            st->nits = st->nits + 1;
            if (qlen > 0 && st->nits >= st->max_iterations) st->done = 1;   // symmetric.rs:326-328
        }
        if (st->safety_rounds_left <= 1) st->done = 1; else st->safety_rounds_left = st->safety_rounds_left - 1;
        if (!st->done && qlen == 0) finish_if_possible(st);
    }
    return qlen;
}

// =============================================================================================================
// Grid-wide assignment + queue compaction (reference src/symmetric.rs:386-463, src/ksparse.rs:229-244).
// One thread per old queue slot; each slot yields 0 or 1 entries of the next queue (loser -> itself,
// winner that evicts -> the evicted owner, winner of a free object or dropped person -> nothing).
// =============================================================================================================
constexpr int kAssignChunk = 2048;   // queue slots one block stages before it reserves output space
#ifndef SLA_ASSIGN_UNROLL
#define SLA_ASSIGN_UNROLL 4
#endif
constexpr int kAssignUnroll = SLA_ASSIGN_UNROLL;   // slots per thread whose dependent loads are issued together

// `first` = 1: first round of a solve behind first_assign_objects_kernel -- the objects have already installed their
// winners (price, owner, cleared word), so a bidder only has to look up whether it is the owner of its object: it
// records the match or goes back into the queue.  Every person is a bidder in that round, so this pass also is the
// initialisation of person_to_object.  (Tried in round 2: the object pass scattering the winners into person_to_object
// so that this pass reads two coalesced words per person -- cfg3 unchanged, cfg5 8 % slower: 13 M scattered 4-byte
// stores into a 64 MB array cost more than 16 M scattered 4-byte loads; gpurun r2x)
__global__ void __launch_bounds__(kWideThreads) assign_wide_kernel(const Params p, const int first) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    const uint32_t cur = h.cur;
    const uint32_t qlen = h.qlen[cur & 1u];
    if (h.done || qlen <= h.tail_max) return;
    const bool identity = h.identity != 0;
    const bool nobody_owns = h.zero_prices != 0;   // first round after init_solve: every object is still free
    const uint32_t pbits = h.pbits;
    const uint32_t* __restrict__ queue = cur ? p.queue[1] : p.queue[0];
    uint32_t* __restrict__ next_queue = cur ? p.queue[0] : p.queue[1];
    uint32_t* next_len = &st->qlen[cur ^ 1u];

    __shared__ uint32_t s_emit[kAssignChunk];
    __shared__ uint32_t s_cnt, s_base;
    const int lane = threadIdx.x & 31;

    // contiguous chunk per block: one global atomicAdd per chunk instead of one per 256 slots
    uint32_t per = (qlen + gridDim.x - 1) / gridDim.x;
    per = ((per + kWideThreads - 1) / kWideThreads) * kWideThreads;
    if (per > (uint32_t)kAssignChunk) per = kAssignChunk;

    for (uint32_t start = blockIdx.x * per; start < qlen; start += gridDim.x * per) {
        const uint32_t stop = (start + per < qlen) ? start + per : qlen;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (uint32_t t0 = start; t0 < stop; t0 += kWideThreads * kAssignUnroll) {
            uint32_t j[kAssignUnroll], i[kAssignUnroll], prev[kAssignUnroll];
            double bid[kAssignUnroll];
            unsigned long long word[kAssignUnroll];
#pragma unroll
            for (int u = 0; u < kAssignUnroll; ++u) {
                const uint32_t q = t0 + u * kWideThreads + threadIdx.x;
                j[u] = SLA_DEV_NONE; i[u] = 0; bid[u] = 0.0;
                if (q < stop) {
                    j[u] = p.slot_obj[q];
                    bid[u] = p.slot_bid[q];
                    i[u] = identity ? q : __ldg(queue + q);
                }
            }
#pragma unroll
            for (int u = 0; u < kAssignUnroll; ++u) {
                word[u] = 0ull; prev[u] = SLA_DEV_NONE;
                if (j[u] != SLA_DEV_NONE) {
                    if (first) prev[u] = __ldcg(p.o2p + j[u]);          // the owner the object pass installed
                    else {
                        word[u] = __ldcg(p.best + j[u]);
                        if (!nobody_owns) prev[u] = __ldcg(p.o2p + j[u]);   // speculative: only the winner uses it
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kAssignUnroll; ++u) {
                uint32_t emit = SLA_DEV_NONE;
                if (first) {
                    const uint32_t q = t0 + u * kWideThreads + threadIdx.x;
                    if (q < stop) {
                        const bool won = j[u] != SLA_DEV_NONE && prev[u] == i[u];
                        p.p2o[i[u]] = won ? j[u] : SLA_DEV_NONE;
                        if (!won && j[u] != SLA_DEV_NONE) emit = i[u];      // dropped persons (no object) leave the auction
                    }
                } else if (j[u] != SLA_DEV_NONE) {
                    const bool won = (bid[u] == bid[u]) && (word[u] == pack_bid(bid[u], i[u], pbits));
                    if (won) {
                        p.prices[j[u]] = bid[u];
                        p.o2p[j[u]] = i[u];
                        p.p2o[i[u]] = j[u];
                        p.best[j[u]] = 0ull;
                        if (prev[u] != SLA_DEV_NONE) { p.p2o[prev[u]] = SLA_DEV_NONE; emit = prev[u]; }
                    } else {
                        emit = i[u];
                    }
                }
                const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
                if (ballot) {
                    uint32_t wbase = 0;
                    if (lane == 0) wbase = atomicAdd(&s_cnt, (uint32_t)__popc(ballot));
                    wbase = __shfl_sync(0xffffffffu, wbase, 0);
                    if (emit != SLA_DEV_NONE) s_emit[wbase + __popc(ballot & ((1u << lane) - 1u))] = emit;
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base = s_cnt ? atomicAdd(next_len, s_cnt) : 0u;
        __syncthreads();
        const uint32_t cnt = s_cnt, gbase = s_base;
        for (uint32_t e = threadIdx.x; e < cnt; e += kWideThreads) next_queue[gbase + e] = s_emit[e];
        __syncthreads();
    }

    // Control step A (round accounting, queue flip, termination tests) by the last block to finish: every block has
    // published its emissions (fence) before it takes its ticket, so the last one sees the complete next queue length.
    // The control block is fetched past the L1 with independent 128-bit loads (other blocks changed it with atomics in
    // the L2), worked on in shared memory and stored back -- one L2 round trip instead of a chain of dependent ones.
    // Only blocks that had a chunk take a ticket (a block without one has written nothing): same-address atomics cost
    // ~8 ns apiece, which is 5 us for a full grid and nothing for the three working blocks of a short round.
    __shared__ uint32_t s_last;
    __shared__ __align__(16) DevState s_state;
    const uint32_t working = min((qlen + per - 1u) / per, gridDim.x);
    if (blockIdx.x >= working) return;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = (atomicAdd(&st->assign_ticket, 1u) == working - 1u) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
        constexpr int kCtlWords = (int)(offsetof(DevState, dbg) / 16);
        static_assert(offsetof(DevState, dbg) % 16 == 0, "the control fields are copied as 128-bit words");
        __threadfence();
        if (threadIdx.x < kCtlWords)
            reinterpret_cast<uint4*>(&s_state)[threadIdx.x] = __ldcg(reinterpret_cast<const uint4*>(st) + threadIdx.x);
        __syncthreads();
        if (threadIdx.x == 0) {
            s_state.assign_ticket = 0u;
            control_after_wide(&s_state);
            s_state.wide_ctl_done = 1u;
        }
        __syncthreads();
        if (threadIdx.x < kCtlWords)
            reinterpret_cast<uint4*>(st)[threadIdx.x] = reinterpret_cast<const uint4*>(&s_state)[threadIdx.x];
    }
}

// =============================================================================================================
// Tail engine: one persistent CTA runs whole Jacobi rounds (bid -> barrier -> assign/compact -> barrier) while
// the queue is short.  Queue and per-slot bids live in shared memory; prices / owners are read through the L1
// (coherent for this CTA's own stores), CSR rows of repeat bidders stay L1-resident.  Conflict resolution:
//   * <= 32 bidders: entirely in shared memory by one warp (no global atomics, no L2 round trips);
//   * more bidders : an open-addressing hash table in shared memory keyed by the object (atomicCAS on the key,
//                    atomicMax on the packed word) -- still no global atomics and no L2 round trips.
// (Control step A of a wide round runs in the last block of assign_wide_kernel.)
// =============================================================================================================
// SPRICES: the object prices are mirrored in dynamic shared memory for the lifetime of the launch (write-through to
// global), which removes k scattered L1 misses per bidder -- the single SM's miss throughput, not latency, is
// what bounds a tail round otherwise (profiles/README.md).  DevState::own_mode says whether the owners (o2p) are
// mirrored as well (1: u32, 2: u16); DevState::tail_cap is the capacity of the queue arrays.  The host sizes the
// dynamic shared memory with tail_smem_bytes() from the same three facts.
template <int LPR, bool SPRICES>
__global__ void __launch_bounds__(kTailThreads, 1) tail_kernel(const Params p) {
    extern __shared__ __align__(16) unsigned char s_dyn[];
    __shared__ __align__(16) uint4 s_xchg[32];   // small rounds: per slot {packed word (2 x u32), object, still-active flag}
    __shared__ uint32_t s_warp_cnt[kTailThreads / 32];
    __shared__ uint32_t s_ctl[2];
    struct SmallFin { uint32_t qlen, nits, hit_limit, claimed; unsigned long long safety, rounds_done, bids_done; };
    __shared__ SmallFin s_fin;
    __shared__ unsigned long long s_arcs;
    __shared__ uint32_t s_dropped;

    // The whole control block is fetched with one cooperative copy, mutated in shared memory by thread 0 and written
    // back at the end: the tail kernel is the only kernel running on the stream, so it owns the block meanwhile.
    __shared__ __align__(16) DevState s_state;
    DevState* const gst = p.st;
    DevState* const st = &s_state;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane32 = tid & 31;
    constexpr int kStateWords = (int)(sizeof(DevState) / 16);
    if (tid < kStateWords) reinterpret_cast<uint4*>(st)[tid] = __ldcg(reinterpret_cast<const uint4*>(gst) + tid);
    __syncthreads();

    if (tid == 0) {
        // control step A of a wide round runs in assign_wide_kernel's last block; a phase action it may have set (Khosla
        // eps-schedule: the queue ran empty) belongs to this super-round's phase kernel and must survive, any older one
        // has been consumed
        if (st->wide_ctl_done) st->wide_ctl_done = 0u; else st->action = ACTION_NONE;
        const uint32_t qlen0 = st->qlen[st->cur];
        const uint32_t run = (!st->done && qlen0 > 0 && qlen0 <= st->tail_own_max) ? 1u : 0u;
        s_ctl[0] = run;
        s_ctl[1] = qlen0;
        s_arcs = 0;
        s_dropped = 0;
        s_fin.claimed = 0u;
    }
    __syncthreads();
    if (!s_ctl[0]) {
        if (tid < kStateWords) reinterpret_cast<uint4*>(gst)[tid] = reinterpret_cast<const uint4*>(st)[tid];
        return;
    }

    uint32_t qlen = s_ctl[1];
    const uint32_t cur = st->cur;
    const bool identity = st->identity != 0;
    const uint32_t algo = st->algo, pbits = st->pbits, max_it = st->max_iterations, sign_flip = st->sign_flip;
    const uint32_t regK = st->regular_k;
    const double eps = st->eps, thr = st->threshold;
    bool zero = (st->zero_prices != 0) && (st->skip_zero != 0);
    uint32_t nits = st->nits;
    unsigned long long safety = st->safety_rounds_left;
    const uint32_t round_cap = st->tail_round_cap;
    const uint32_t n_cols = st->n_cols, own_mode = st->own_mode, cap = st->tail_cap;

    // ---- carve the dynamic shared memory (same arithmetic as tail_smem_bytes) ----
    const TailSmemLayout lay = tail_smem_layout(SPRICES, own_mode, n_cols, cap);
    double* const s_prices = reinterpret_cast<double*>(s_dyn);
    OwnerView own;
    own.g = p.o2p;
    own.s32 = (own_mode == 1u) ? reinterpret_cast<uint32_t*>(s_dyn + lay.owners) : nullptr;
    own.s16 = (own_mode == 2u) ? reinterpret_cast<uint16_t*>(s_dyn + lay.owners) : nullptr;
    unsigned long long* const h_word = reinterpret_cast<unsigned long long*>(s_dyn + lay.hash_words);
    uint32_t* const h_key = reinterpret_cast<uint32_t*>(s_dyn + lay.hash_keys);
    double* const s_bid = reinterpret_cast<double*>(s_dyn + lay.bid);
    uint32_t* const s_queue0 = reinterpret_cast<uint32_t*>(s_dyn + lay.queue0);
    uint32_t* const s_queue1 = s_queue0 + cap;
    uint32_t* const s_obj = s_queue1 + cap;
    uint32_t* const s_prev = s_obj + cap;
    uint32_t* const s_hslot = s_prev + cap;
    const uint32_t hslots = 2u * cap, hshift = 32u - (uint32_t)__ffs((int)hslots) + 1u;   // hslots is a power of two

    uint32_t* gqueue = cur ? p.queue[1] : p.queue[0];
    for (uint32_t q = tid; q < qlen; q += kTailThreads) s_queue0[q] = identity ? q : gqueue[q];
    if (SPRICES)
        for (uint32_t j = tid; j < n_cols; j += kTailThreads) s_prices[j] = zero ? 0.0 : p.prices[j];
    if (own.s32)
        for (uint32_t j = tid; j < n_cols; j += kTailThreads) own.s32[j] = p.o2p[j];
    if (own.s16)
        for (uint32_t j = tid; j < n_cols; j += kTailThreads) own.s16[j] = (uint16_t)p.o2p[j];
    for (uint32_t h = tid; h < hslots; h += kTailThreads) { h_key[h] = SLA_DEV_NONE; h_word[h] = 0ull; }
    __syncthreads();
    constexpr int kPriceMode = SPRICES ? PRICE_SMEM : PRICE_CA;
    const double* price_src = SPRICES ? s_prices : p.prices;

    constexpr int NGROUPS = kTailThreads / LPR;
    const int lane = tid % LPR;
    const uint32_t group = tid / LPR;
    uint32_t buf = 0;
    unsigned long long my_arcs = 0, rounds_done = 0, bids_done = 0;
    uint32_t my_dropped = 0;
    bool hit_limit = false;

#ifdef SLA_TAIL_TIMING
    // Development instrumentation: cycle stamps of the busiest warp of the small rounds.  Each stamp is taken behind an
    // instruction that consumes `dep`, so that it is not issued before that value has arrived.
    long long tk0 = clock64(), tk1 = 0, tkv[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long act_rounds = 0;
    __shared__ unsigned long long s_busiest;
    if (tid == 0) s_busiest = 0ull;
#define TK(i, dep) do { unsigned long long t_; asm volatile("{\n .reg .pred p;\n setp.eq.u32 p, %1, 0x7ffffff1;\n @p trap;\n mov.u64 %0, %%clock64;\n}" \
        : "=l"(t_) : "r"((uint32_t)(dep)) : "memory"); tkv[i] += (long long)t_ - tk1; tk1 = (long long)t_; } while (0)
#else
#define TK(i, dep) do { } while (0)
#endif
    while (true) {
        if (qlen <= kTailSlots) {
            // =================================================================================================
            // Small rounds (<= one bidder per warp): slot-stable, one warp per slot.  Warp q keeps serving "slot q":
            // after a round the slot holds the loser itself, the owner it evicted, or nothing -- no queue compaction,
            // no single-warp assignment phase.  Warps whose slot is empty SLEEP at barrier 0 until the engine is done;
            // the bidders of a round synchronise among themselves on barrier 1 (bar.sync 1, 32 * #bidders), so an
            // empty slot costs neither issue slots nor barrier latency.  The row of the slot's person lives in
            // registers (rows of <= 32 RPL arcs); the row of the owner that would be evicted is fetched as soon as the
            // choice is known, so that its L2 latency hides behind the bid arithmetic, both barriers and the conflict
            // resolution.  Conflicts: every bidder compares its own packed word with the words of all slots (one
            // vote).  A Jacobi round is independent of the order of its bidders, so the results equal the compacting
            // formulation bit for bit.
            // =================================================================================================
            constexpr int RPL = (LPR <= 8) ? 1 : ((LPR == 16) ? 2 : 4);
            const uint32_t* sq0 = buf ? s_queue1 : s_queue0;
            uint32_t mask = (qlen == 32u) ? 0xffffffffu : ((1u << qlen) - 1u);   // qlen <= kTailSlots <= 32
            bool active = (uint32_t)warp < qlen;
            uint32_t person = active ? sq0[warp] : 0u;
            __syncthreads();                             // the queue buffer is rewritten at the end
            if (active) {
                // Shared-space addresses, computed once.  `opq` is a run-time zero the compiler cannot see through: without
                // it ptxas rematerialises the bases inside the loop (S2UR SR_CgaCtaId / SR_SWINHI chains per access).
                const uint32_t opq = st->hot_pad[0];
                const uint32_t a_xchg = smem_base_u32(s_xchg) + opq;
                const uint32_t a_mine = a_xchg + 16u * (uint32_t)warp, a_lane = a_xchg + 16u * (uint32_t)lane32;
                const uint32_t a_prices = smem_base_u32(s_prices) + opq;
                const uint32_t a_own = smem_base_u32(s_dyn + lay.owners) + opq;
                const uint32_t* const cols_lane = p.cols + lane32;
                const double* const vals_lane = p.vals + lane32;
                // Rounds this engine may still run: the limit that ends the solve (max_iterations of the Forward solver,
                // symmetric.rs:326-328, or the safety budget) and the hand-back cap of one launch, as one countdown.
                const unsigned long long lim_iter = (algo == ALGO_FORWARD) ? (nits < max_it ? (unsigned long long)(max_it - nits) : 1ull) : ~0ull;
                const unsigned long long lim_hit = lim_iter < safety ? lim_iter : (safety ? safety : 1ull);
                const unsigned long long lim_cap = rounds_done < round_cap ? (unsigned long long)round_cap - rounds_done : 1ull;
                unsigned long long lim = lim_hit < lim_cap ? lim_hit : lim_cap;
                const bool budget_ends_solve = lim_hit <= lim_cap && lim <= 0x7fffffffull;   // running out of budget == hitting a limit
                if (lim > 0x7fffffffull) lim = 0x7fffffffull;
                const uint32_t budget = (uint32_t)lim;
                uint32_t left = budget, bids_local = 0u, arcs_local = 0u;
                RowRegs<RPL> row;
                row_extents<RPL>(row, p.row_ptr, regK, person);
                row_load<RPL>(row, cols_lane, vals_lane, lane32);
                bool next_active = false;
                while (true) {
                    const uint32_t nbid = (uint32_t)__popc(mask);
                    // Solo chain: this warp is the only bidder left, and the set of bidders never grows, so every
                    // remaining round of this launch has exactly one bid (two thirds of all rounds of the eps-scaled
                    // solvers are like that: an eviction chain running to its end).  Nothing to exchange, nobody to
                    // wait for: no packed word, no exchange record, no barriers, no vote.
                    const bool solo = nbid == 1u;
#ifdef SLA_TAIL_TIMING
                    if (tk1 == 0) tk1 = clock64();
                    act_rounds += 1;
#endif
                    // ---- bid ----
                    uint32_t obj = SLA_DEV_NONE, prev = SLA_DEV_NONE;
                    unsigned long long word = 0ull;
                    double bid = 0.0;
                    RowRegs<RPL> nrow;
                    nrow.a = 0u; nrow.len = 0u;
                    WarpChoice c;
                    if (row.len <= 32u * RPL) {
                        const LaneTop2 t = SPRICES ? lane_scan_regs_smem<RPL>(row, a_prices, sign_flip, lane32)
                                                   : lane_scan_regs<RPL, PRICE_CA>(row, p.prices, sign_flip, lane32);
                        TK(0, t.k1);
                        if (own_mode) c = warp_choice_finish<OWN_NONE>(t, nullptr);
                        else c = warp_choice_finish<OWN_GLOBAL>(t, p.o2p);
                    } else {
                        if (own_mode) c = warp_bid_scan<kPriceMode, OWN_NONE>(p.cols, p.vals, price_src, nullptr, row.a, row.a + row.len, sign_flip, lane32);
                        else c = warp_bid_scan<kPriceMode, OWN_GLOBAL>(p.cols, p.vals, price_src, p.o2p, row.a, row.a + row.len, sign_flip, lane32);
                    }
                    TK(1, c.col);
                    // owner of the chosen object (c.col is 0 for a row without usable arc, like the reference's start object)
                    if (own_mode == 1u) c.owner = lds_u32(a_own + 4u * c.col);
                    else if (own_mode == 2u) { const uint32_t o16 = lds_u16(a_own + 2u * c.col); c.owner = (o16 == 0xFFFFu) ? SLA_DEV_NONE : o16; }
                    else if (c.pos == SLA_DEV_NONE) c.owner = ld_ca_u32(p.o2p);
                    prev = c.owner;
                    TK(2, prev);
                    if (prev != SLA_DEV_NONE) {          // whoever wins this object evicts `prev`: fetch its row now
                        row_extents<RPL>(nrow, p.row_ptr, regK, prev);
                        row_load<RPL>(nrow, cols_lane, vals_lane, lane32);
                    }
                    const Bid r = make_bid_choice<kPriceMode>(c, algo, eps, thr, price_src);
                    arcs_local += row.len;
                    if (r.dropped) {
                        if (lane32 == 0) my_dropped += 1;
                    } else {
                        obj = r.obj;
                        bid = r.bid;
                        if (solo) word = (r.bid == r.bid) ? 1ull : 0ull;                   // a lone bid wins unless it is NaN
                        else word = (r.bid == r.bid) ? pack_bid(r.bid, person, pbits) : 0ull;   // NaN never bids
                    }
                    TK(3, (uint32_t)word);
                    const bool lv = ((mask >> lane32) & 1u) != 0u;
                    if (!solo) {
                        if (lane32 == 0) { sts_u64(a_mine, word); sts_u32(a_mine + 8u, obj); }
                        TK(4, 0);
                        named_bar_sync(32u * nbid);
                        TK(5, 0);
                    }
                    // ---- resolve + assign ----
                    next_active = false;
                    if (obj != SLA_DEV_NONE) {
                        bool lost = false;
                        if (!solo) {
                            uint4 rec = make_uint4(0u, 0u, SLA_DEV_NONE, 0u);
                            if (lv) rec = lds_u128(a_lane);  // {word lo, word hi, object, flag of the previous round}
                            const unsigned long long rw = ((unsigned long long)rec.y << 32) | rec.x;
                            lost = __any_sync(0xffffffffu, rec.z == obj && rw > word);
                        }
                        TK(6, lost);
                        if (word != 0ull && !lost) {     // word 0 == NaN bid: never wins
                            if (lane32 == 0) {
                                p.prices[obj] = bid;
                                if (SPRICES) sts_f64(a_prices + 8u * obj, bid);
                                p.o2p[obj] = person;
                                if (own_mode == 1u) sts_u32(a_own + 4u * obj, person);
                                else if (own_mode == 2u) sts_u16(a_own + 2u * obj, person);
                                p.p2o[person] = obj;
                                if (prev != SLA_DEV_NONE) p.p2o[prev] = SLA_DEV_NONE;
                            }
                            if (prev != SLA_DEV_NONE) { person = prev; row = nrow; next_active = true; }
                        } else {
                            next_active = true;          // lost: bids again
                        }
                        TK(7, row.c[0]);
                    }
                    uint32_t nmask;
                    if (solo) {
                        nmask = next_active ? mask : 0u;
                        __syncwarp();                    // lane 0's mirror stores before the next round's gather
                    } else {
                        if (lane32 == 0) sts_u32(a_mine + 12u, next_active ? 1u : 0u);
                        TK(8, 0);
                        named_bar_sync(32u * nbid);
                        TK(9, 0);
                        nmask = __ballot_sync(0xffffffffu, lv && lds_u32(a_lane + 12u) != 0u);
                    }
                    bids_local += nbid;
                    mask = nmask;
                    left -= 1u;
                    TK(10, nmask);
                    if (nmask == 0u || left == 0u || !next_active) break;   // done / out of budget / this slot is empty from now on
                }
                if (mask == 0u || left == 0u) {
                    // every bidder of the last round holds the same complete counters: one of them publishes the totals
                    const uint32_t r = budget - left;                       // rounds this engine ran
                    const bool hit = mask != 0u && left == 0u && budget_ends_solve;
                    if (lane32 == 0 && s_fin_claim(&s_fin.claimed)) {
                        s_fin.qlen = (uint32_t)__popc(mask);
                        s_fin.nits = (algo == ALGO_FORWARD) ? nits + r : nits;
                        s_fin.hit_limit = hit ? 1u : 0u;
                        // the safety budget is charged for every round that neither emptied the queue nor hit a limit
                        s_fin.safety = safety - (unsigned long long)(r - 1u) - ((mask != 0u && !hit) ? 1ull : 0ull);
                        s_fin.rounds_done = rounds_done + r;
                        s_fin.bids_done = bids_done + bids_local;
                    }
                    // compact the surviving slots into the queue buffer for the write-back below
                    if (next_active && lane32 == 0) (buf ? s_queue1 : s_queue0)[__popc(mask & ((1u << warp) - 1u))] = person;
                }
                if (lane32 == 0) my_arcs += (unsigned long long)arcs_local;
            }
            __syncthreads();
            qlen = s_fin.qlen; nits = s_fin.nits; hit_limit = s_fin.hit_limit != 0u;
            safety = s_fin.safety; rounds_done = s_fin.rounds_done; bids_done = s_fin.bids_done;
            break;
        }
        // ---- bidding phase (33 .. cap bidders): one group of LPR lanes per bidder ----
        const uint32_t* sq = buf ? s_queue1 : s_queue0;
        for (uint32_t base = 0; base < qlen; base += NGROUPS) {
            // warps none of whose groups has a bidder leave (warp-uniform, so the full-mask shuffles stay legal):
            // idle warps must not burn issue slots on the reduction while one or two warps do the real work
            if (base + (uint32_t)(warp * 32) / LPR >= qlen) break;
            const uint32_t q = base + group;
            const bool valid = q < qlen;
            uint32_t i = 0, a = 0, b = 0;
            if (valid) {
                i = sq[q];
                if (regK) { a = i * regK; b = a + regK; }
                else { a = __ldg(p.row_ptr + i); b = __ldg(p.row_ptr + i + 1); }
            }
            Choice c;
            choice_init(c);
            if (zero) scan_row<LPR, PRICE_ZERO, false>(c, p.cols, p.vals, price_src, a, b, sign_flip, lane);
            else      scan_row<LPR, kPriceMode, false>(c, p.cols, p.vals, price_src, a, b, sign_flip, lane);
            // speculative: the owner of each lane's local best column, issued before the reduction so that its latency
            // hides behind the shuffles (nobody owns anything while all prices are still zero)
            if (!zero && c.pos != SLA_DEV_NONE) c.aux = own.load(c.col);
            choice_group_reduce<LPR>(c);
            if (valid && lane == 0 && !zero && c.pos == SLA_DEV_NONE) c.aux = own.load(0u);   // row without usable arc -> object 0
            if (valid && lane == 0) {
                const Bid r = zero ? make_bid<PRICE_ZERO>(c, algo, eps, thr, price_src)
                                   : make_bid<kPriceMode>(c, algo, eps, thr, price_src);
                my_arcs += (unsigned long long)(b - a);
                if (r.dropped) {
                    s_obj[q] = SLA_DEV_NONE;
                    my_dropped += 1;
                } else {
                    s_obj[q] = r.obj;
                    s_bid[q] = r.bid;
                    s_prev[q] = c.aux;                      // owner of r.obj (only the winner uses it)
                    if (r.bid == r.bid) {                   // NaN never bids
                        uint32_t h = (r.obj * 2654435761u) >> hshift;
                        while (true) {
                            const uint32_t old = atomicCAS(h_key + h, SLA_DEV_NONE, r.obj);
                            if (old == SLA_DEV_NONE || old == r.obj) break;
                            h = (h + 1u) & (hslots - 1u);
                        }
                        atomicMax(h_word + h, pack_bid(r.bid, i, pbits));
                        s_hslot[q] = h;
                    }
                }
            }
        }
        __syncthreads();

        // ---- assignment phase + deterministic compaction into the other smem queue ----
        uint32_t* nq = buf ? s_queue0 : s_queue1;
        uint32_t out = 0;
        for (uint32_t base = 0; base < qlen; base += kTailThreads) {
            const uint32_t q = base + tid;
            uint32_t emit = SLA_DEV_NONE;
            if (q < qlen) {
                const uint32_t j = s_obj[q];
                if (j != SLA_DEV_NONE) {
                    const uint32_t i = sq[q];
                    const double bid = s_bid[q];
                    const uint32_t prev = s_prev[q];
                    const bool won = (bid == bid) && (h_word[s_hslot[q]] == pack_bid(bid, i, pbits));
                    if (won) {
                        p.prices[j] = bid;
                        if (SPRICES) s_prices[j] = bid;
                        p.o2p[j] = i;
                        own.store_mirror(j, i);
                        p.p2o[i] = j;
                        if (prev != SLA_DEV_NONE) { p.p2o[prev] = SLA_DEV_NONE; emit = prev; }
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
            for (int w = 0; w < kTailThreads / 32; ++w) {
                const uint32_t cnt = s_warp_cnt[w];
                off += (w < warp) ? cnt : 0u;
                total += cnt;
            }
            if (emit != SLA_DEV_NONE) nq[out + off + __popc(ballot & ((1u << lane32) - 1u))] = emit;
            out += total;
            __syncthreads();
        }
        // every word has been compared: empty the table entries this round used
        for (uint32_t q = tid; q < qlen; q += kTailThreads) {
            if (s_obj[q] != SLA_DEV_NONE && s_bid[q] == s_bid[q]) {
                const uint32_t h = s_hslot[q];
                h_key[h] = SLA_DEV_NONE;
                h_word[h] = 0ull;
            }
        }
        __syncthreads();

        bids_done += qlen;
        rounds_done += 1;
        qlen = out;
        buf ^= 1u;
        zero = false;
        if (algo == ALGO_FORWARD) nits += 1;
        if (qlen == 0) break;
        if (algo == ALGO_FORWARD && nits >= max_it) { hit_limit = true; break; }   // symmetric.rs:326-328
        if (safety <= 1) { hit_limit = true; break; }
        safety -= 1;
        if (rounds_done >= round_cap) break;   // hand control back; the next super-round continues
    }

    // ---- write the state back ----
    if (my_arcs) atomicAdd(&s_arcs, my_arcs);
    if (my_dropped) atomicAdd(&s_dropped, my_dropped);
    for (uint32_t q = tid; q < qlen; q += kTailThreads) gqueue[q] = (buf ? s_queue1 : s_queue0)[q];
    __syncthreads();
    if (tid == 0) {
        st->rounds += rounds_done;
        st->tail_rounds += rounds_done;
        st->bids += bids_done;
        st->bid_arcs += s_arcs;
        st->dropped += s_dropped;
        st->qlen[cur] = qlen;
        st->identity = 0;
        st->zero_prices = 0;
        st->nits = nits;
        st->safety_rounds_left = safety;
        if (hit_limit) st->done = 1;
        else if (qlen == 0) finish_if_possible(st);
#ifdef SLA_TAIL_TIMING
        st->dbg[0] += (unsigned long long)(clock64() - tk0);
        st->dbg[6] += rounds_done;
#endif
    }
#ifdef SLA_TAIL_TIMING
    if (lane32 == 0) atomicMax(&s_busiest, (act_rounds << 5) | (unsigned long long)warp);
    __syncthreads();
    if (lane32 == 0 && (s_busiest & 31ull) == (unsigned long long)warp && act_rounds) {
        st->dbg[7] += act_rounds;
        for (int i = 0; i < 12; ++i) st->dbg[8 + i] += (unsigned long long)tkv[i];
    }
#endif
    __syncthreads();
    if (tid < kStateWords) reinterpret_cast<uint4*>(gst)[tid] = reinterpret_cast<const uint4*>(st)[tid];
#undef TK
}

// =============================================================================================================
// Cluster engine: ONE thread-block cluster of kMidCtas CTAs runs whole Jacobi rounds while the queue holds between
// DevState::tail_own_max and DevState::tail_max (<= kMidMax) bidders.  It exists for instances whose prices do not fit
// the tail engine's shared memory (large M): there a round of a few thousand bidders costs two grid-wide launches
// (~14 us in the solve graph, almost all of it launch and drain) or, in the single CTA, several serial passes of
// dependent global loads (~5 us).  Here the queue is SPLIT over the CTAs and never leaves their shared memory:
//   bid      every CTA scans the rows of its own bidders (prices past the L1, they change between the rounds of this
//            launch) and posts the packed bid words with atomicMax into the same global words the wide kernels use;
//   barrier  cluster-wide (release / acquire);
//   resolve  one thread per local slot compares its word with the winner's; winners install price / owner / assignment
//            and clear the word; the loser itself or the evicted owner goes into the CTA's OWN next queue (a slot yields
//            at most one entry, so a local queue never grows); the local lengths are exchanged through distributed
//            shared memory;
//   barrier  cluster-wide; every CTA now knows the total and takes the same decision.
// No global queue, no global counter, no host: two cluster barriers (~0.2 us each) per round.  A Jacobi round does not
// depend on the order or the placement of its bidders, so the results equal those of the other engines bit for bit.
// The engine hands back when the queue is short enough for the single-CTA engine (its slot-stable rounds cost ~1 us),
// empty, or a limit is reached; the local queues are then concatenated into the global queue buffer.
// =============================================================================================================
template <int LPR>
__global__ void __cluster_dims__(kMidCtas, 1, 1) __launch_bounds__(kMidThreads, 1) mid_kernel(const Params p) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ uint32_t s_queue[2][kMidLocalCap];
    __shared__ uint32_t s_obj[kMidLocalCap];
    __shared__ __align__(8) double s_bid[kMidLocalCap];
    __shared__ uint32_t s_counts[2][kMidCtas];          // by round parity: next-queue length of every CTA of the cluster
    __shared__ unsigned long long s_arcs_all[kMidCtas]; // rank 0's copy is the one that is read
    __shared__ uint32_t s_dropped_all[kMidCtas];
    __shared__ uint32_t s_len;
    __shared__ unsigned long long s_arcs;
    __shared__ uint32_t s_dropped;
    __shared__ __align__(16) DevState s_state;

    DevState* const gst = p.st;
    const HotState h = load_hot(gst);
    const uint32_t cur = h.cur & 1u;
    const uint32_t qlen0 = h.qlen[cur];
    const uint32_t own_max = gst->tail_own_max;
    // the same values in every CTA: either the whole cluster leaves here or none of it
    if (h.done || qlen0 <= own_max || qlen0 > h.tail_max || qlen0 > kMidMax) return;

    const int tid = threadIdx.x, warp = tid >> 5, lane32 = tid & 31;
    constexpr int G = kMidThreads / LPR;               // bidders one CTA scans per pass
    const int lane = tid % LPR;
    const uint32_t group = tid / LPR;
    const uint32_t rank = cluster.block_rank();
    const bool identity = h.identity != 0;
    const uint32_t algo = h.algo, pbits = h.pbits, sign_flip = h.sign_flip, regK = h.regular_k;
    const double eps = h.eps, thr = h.threshold;
    const uint32_t max_it = gst->max_iterations, round_cap = gst->tail_round_cap;
    uint32_t nits = gst->nits;
    unsigned long long safety = gst->safety_rounds_left;
    uint32_t* const gqueue = cur ? p.queue[1] : p.queue[0];

    // contiguous share of the global queue
    const uint32_t share = (qlen0 + kMidCtas - 1u) / kMidCtas;
    const uint32_t begin = rank * share < qlen0 ? rank * share : qlen0;
    uint32_t n = (begin + share < qlen0 ? begin + share : qlen0) - begin;
    for (uint32_t l = tid; l < n; l += kMidThreads) s_queue[0][l] = identity ? begin + l : __ldcg(gqueue + begin + l);
    if (tid == 0) { s_arcs = 0ull; s_dropped = 0u; s_len = 0u; }
    __syncthreads();

    uint32_t total = qlen0, buf = 0u;
    unsigned long long rounds_done = 0ull, bids_done = 0ull, my_arcs = 0ull;
    uint32_t my_dropped = 0u;
    bool hit_limit = false;
#ifdef SLA_MID_TIMING
    // development instrumentation: cycle stamps of thread 0 of cluster rank 0 (in-order issue: a stamp behind an
    // instruction that consumes loaded data is taken after the data has arrived)
    long long mt[6] = {0, 0, 0, 0, 0, 0}, mlast, mround;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(mlast) :: "memory");
    mround = mlast;
#define MT(i) do { long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory"); mt[i] += t_ - mlast; mlast = t_; } while (0)
#else
#define MT(i) do { } while (0)
#endif
    while (true) {
        // ---- bid: one group of LPR lanes per local bidder ----
        const uint32_t* sq = s_queue[buf];
        for (uint32_t base = 0; base < n; base += G) {
            if (base + (uint32_t)(warp * 32) / LPR >= n) break;   // warp-uniform: the full-mask shuffles stay legal
            const uint32_t q = base + group;
            const bool valid = q < n;
            uint32_t i = 0, a = 0, b = 0;
            if (valid) {
                i = sq[q];
                if (regK) { a = i * regK; b = a + regK; }
                else { a = __ldg(p.row_ptr + i); b = __ldg(p.row_ptr + i + 1); }
            }
            Choice c;
            choice_init(c);
            scan_row<LPR, PRICE_CG, false>(c, p.cols, p.vals, p.prices, a, b, sign_flip, lane);
            choice_group_reduce<LPR>(c);
            if (valid && lane == 0) {
                const Bid r = make_bid<PRICE_CG>(c, algo, eps, thr, p.prices);
                my_arcs += (unsigned long long)(b - a);
                if (r.dropped) {
                    s_obj[q] = SLA_DEV_NONE;
                    my_dropped += 1u;
                } else {
                    s_obj[q] = r.obj;
                    s_bid[q] = r.bid;
                    if (r.bid == r.bid) atomicMax(p.best + r.obj, pack_bid(r.bid, i, pbits));   // NaN never bids
                }
            }
        }
        MT(0);
        cluster.sync();                                   // every bid word of the round is in the L2
        MT(1);

        // ---- resolve: one thread per local slot ----
        uint32_t* nq = s_queue[buf ^ 1u];
        for (uint32_t base = 0; base < n; base += kMidThreads) {
            const uint32_t q = base + tid;
            uint32_t emit = SLA_DEV_NONE;
            if (q < n) {
                const uint32_t j = s_obj[q];
                if (j != SLA_DEV_NONE) {
                    const uint32_t i = sq[q];
                    const double bid = s_bid[q];
                    const unsigned long long word = __ldcg(p.best + j);
                    const uint32_t prev = __ldcg(p.o2p + j);      // only the winner uses it (nobody else writes it this round)
                    const bool won = (bid == bid) && (word == pack_bid(bid, i, pbits));
                    if (won) {
                        p.prices[j] = bid;
                        p.o2p[j] = i;
                        p.p2o[i] = j;
                        p.best[j] = 0ull;                         // a loser that reads 0 instead of the winning word has lost all the same
                        if (prev != SLA_DEV_NONE) {
                            p.p2o[prev] = SLA_DEV_NONE;
                            emit = prev;
                            if (regK && regK <= 64u) {            // the evicted owner bids next round: pull its row towards the L2
                                const char* rc = reinterpret_cast<const char*>(p.cols + (size_t)prev * regK);
                                const char* rv = reinterpret_cast<const char*>(p.vals + (size_t)prev * regK);
                                for (uint32_t off = 0; off < regK * 4u; off += 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(rc + off));
                                for (uint32_t off = 0; off < regK * 8u; off += 128u) asm volatile("prefetch.global.L2 [%0];" :: "l"(rv + off));
                            }
                        }
                    } else {
                        emit = i;
                    }
                }
            }
            const uint32_t ballot = __ballot_sync(0xffffffffu, emit != SLA_DEV_NONE);
            if (ballot) {
                uint32_t wbase = 0;
                if (lane32 == 0) wbase = atomicAdd(&s_len, (uint32_t)__popc(ballot));
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                if (emit != SLA_DEV_NONE) nq[wbase + __popc(ballot & ((1u << lane32) - 1u))] = emit;
            }
        }
        MT(2);
        __syncthreads();
        const uint32_t n_next = s_len;
        const uint32_t par = (uint32_t)(rounds_done & 1ull);
        if (tid < kMidCtas) *cluster.map_shared_rank(&s_counts[par][rank], tid) = n_next;   // my length, into every CTA
        MT(3);
        cluster.sync();                                   // prices / owners / cleared words / lengths of the round are visible
        MT(4);
        uint32_t total_next = 0;
#pragma unroll
        for (int r = 0; r < kMidCtas; ++r) total_next += s_counts[par][r];
        if (tid == 0) s_len = 0u;                         // next use is behind the next round's first barrier

#ifdef SLA_MID_TIMING
        if (rank == 0u && tid == 0 && rounds_done < 8ull) {
            long long t_; asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_) :: "memory");
            gst->dbg[8 + 2 * rounds_done] = (unsigned long long)(t_ - mround);       // cycles of this round
            gst->dbg[9 + 2 * rounds_done] = total;                                     // its bidders
            mround = t_;
        }
#endif
        bids_done += total;
        rounds_done += 1ull;
        total = total_next;
        n = n_next;
        buf ^= 1u;
        if (algo == ALGO_FORWARD) nits += 1u;
        if (total == 0u) break;
        if (algo == ALGO_FORWARD && nits >= max_it) { hit_limit = true; break; }   // symmetric.rs:326-328
        if (safety <= 1ull) { hit_limit = true; break; }
        safety -= 1ull;
        if (total <= own_max) break;                      // short enough for the single-CTA engine
        if (rounds_done >= round_cap) break;              // hand control back; the next super-round continues
    }

#ifdef SLA_MID_TIMING
    if (rank == 0u && tid == 0) {
        for (int i = 0; i < 5; ++i) atomicAdd(&gst->dbg[i], (unsigned long long)mt[i]);
        atomicAdd(&gst->dbg[6], rounds_done);
    }
#endif
#undef MT
    // ---- hand the queue back: local queues concatenated in rank order ----
    const uint32_t lpar = (uint32_t)((rounds_done - 1ull) & 1ull);
    uint32_t off = 0;
#pragma unroll
    for (int r = 0; r < kMidCtas; ++r) off += ((uint32_t)r < rank) ? s_counts[lpar][r] : 0u;
    for (uint32_t l = tid; l < n; l += kMidThreads) gqueue[off + l] = s_queue[buf][l];
    if (my_arcs) atomicAdd(&s_arcs, my_arcs);
    if (my_dropped) atomicAdd(&s_dropped, my_dropped);
    __syncthreads();
    if (tid == 0) {
        *cluster.map_shared_rank(&s_arcs_all[rank], 0) = s_arcs;
        *cluster.map_shared_rank(&s_dropped_all[rank], 0) = s_dropped;
    }
    cluster.sync();
    if (rank != 0u) return;

    // ---- rank 0: the control block, fetched past the L1 in one go, updated in shared memory, stored back ----
    constexpr int kCtlWords = (int)(offsetof(DevState, dbg) / 16);
    if (tid < kCtlWords) reinterpret_cast<uint4*>(&s_state)[tid] = __ldcg(reinterpret_cast<const uint4*>(gst) + tid);
    __syncthreads();
    if (tid == 0) {
        DevState* const st = &s_state;
        unsigned long long arcs = 0ull;
        uint32_t dropped = 0u;
        for (int r = 0; r < kMidCtas; ++r) { arcs += s_arcs_all[r]; dropped += s_dropped_all[r]; }
        // control step A of a wide round of this super-round may have left a phase action for the phase kernel: it stays
        // (the queue was not empty then, so there is none in practice); an older one has been consumed.  The tail engine
        // that follows keeps whatever this engine decides (wide_ctl_done).
        if (!st->wide_ctl_done) st->action = ACTION_NONE;
        st->wide_ctl_done = 1u;
        st->rounds += rounds_done;
        st->tail_rounds += rounds_done;
        st->cluster_rounds += (uint32_t)rounds_done;
        st->bids += bids_done;
        st->bid_arcs += arcs;
        st->dropped += dropped;
        st->qlen[cur] = total;
        st->identity = 0u;
        st->zero_prices = 0u;
        st->nits = nits;
        st->safety_rounds_left = safety;
        if (hit_limit) st->done = 1u;
        else if (total == 0u) finish_if_possible(st);
    }
    __syncthreads();
    if (tid < kCtlWords) reinterpret_cast<uint4*>(gst)[tid] = reinterpret_cast<const uint4*>(&s_state)[tid];
}

// =============================================================================================================
// Forward: eps-complementary-slackness check (reference src/solver.rs:154-189), then control step B in the
// last block to finish (reference src/symmetric.rs:278-328): optimal / give up / start the next eps phase.
// =============================================================================================================
template <int LPR>
__global__ void __launch_bounds__(kWideThreads) ecs_kernel(const Params p) {
    DevState* st = p.st;
    const HotState h = load_hot(st);
    if (h.done || h.algo != ALGO_FORWARD || h.start_opt || h.qlen[h.cur & 1u] != 0) return;
    const double eps = h.target_eps, tol = h.tol;
    const uint32_t n_rows = h.n_rows, n_cols = h.n_cols, sign_flip = h.sign_flip;

    constexpr int GROUPS_PER_BLOCK = kWideThreads / LPR;
    const int lane = threadIdx.x % LPR;
    const uint32_t group = blockIdx.x * GROUPS_PER_BLOCK + threadIdx.x / LPR;
    const uint32_t ngroups = gridDim.x * GROUPS_PER_BLOCK;
    bool violated = false;

    for (uint32_t base = 0; base < n_rows; base += ngroups) {
        const uint32_t i = base + group;
        const bool valid = i < n_rows;
        uint32_t a = 0, b = 0, j = 0;
        if (valid) {
            a = __ldg(p.row_ptr + i);
            b = __ldg(p.row_ptr + i + 1);
            j = p.p2o[i];
        }
        // pass 1: value of the LAST arc of the row that points at the chosen object (solver.rs:163-169)
        uint32_t cpos = 0;   // position + 1, 0 = none
        double cval = neg_inf();
        for (uint32_t g = a + lane; g < b; g += LPR) {
            if (__ldg(p.cols + g) == j) { cpos = g + 1u; cval = __ldg(p.vals + g); }
        }
#pragma unroll
        for (int m = LPR / 2; m >= 1; m >>= 1) {
            const uint32_t op = __shfl_xor_sync(0xffffffffu, cpos, m);
            const double ov = __shfl_xor_sync(0xffffffffu, cval, m);
            if (op > cpos) { cpos = op; cval = ov; }
        }
        if (valid) {
            if (j >= n_cols) {
                violated = true;   // unassigned person: cannot be a complete eps-CS solution
            } else {
                const double chosen = (cpos == 0u) ? neg_inf()
                                                   : __hiloint2double(__double2hiint(cval) ^ (int)sign_flip, __double2loint(cval));
                const double lhs = chosen - __ldg(p.prices + j) + tol;
                // pass 2 (solver.rs:177-186)
                for (uint32_t g = a + lane; g < b; g += LPR) {
                    const double raw = __ldg(p.vals + g);
                    const double v = __hiloint2double(__double2hiint(raw) ^ (int)sign_flip, __double2loint(raw));
                    const double pk = __ldg(p.prices + __ldg(p.cols + g));
                    if (lhs < v - pk - eps) violated = true;
                }
            }
        }
    }
    if (violated) st->ecs_violated = 1;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t ticket = atomicAdd(&st->ecs_ticket, 1u);
        if (ticket == gridDim.x - 1) {
            __threadfence();
            const uint32_t bad = *((volatile uint32_t*)&st->ecs_violated);
            if (!bad) {
                st->optimal = 1;
                st->done = 1;
            } else if (st->eps < st->target_eps) {
                st->done = 1;                                   // symmetric.rs:291-294
            } else {
                st->eps *= 0.15;                                // REDUCTION_FACTOR, symmetric.rs:189, 296
                st->nreductions += 1;
                st->action = ACTION_RESET;
                st->qlen[st->cur] = n_rows;
                st->identity = 1;
                if (st->nits >= st->max_iterations) st->done = 1;   // symmetric.rs:326-328 (after the reset)
            }
            st->ecs_violated = 0;
            st->ecs_ticket = 0;
        }
    }
}

// Phase restart: wipe both assignment vectors, keep prices (reference src/symmetric.rs:299-321; also the phases of
// the Khosla eps-schedule).  ACTION_RESET_ALL zeroes the prices as well (Khosla starting over without a schedule).
__global__ void __launch_bounds__(kWideThreads) phase_apply_kernel(const Params p) {
    const HotState h = load_hot(p.st);
    if (h.action != ACTION_RESET && h.action != ACTION_RESET_ALL) return;
    const bool all = h.action == ACTION_RESET_ALL;
    const uint32_t n_rows = h.n_rows, n_cols = h.n_cols;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n_rows; i += stride) p.p2o[i] = SLA_DEV_NONE;
    for (uint32_t j = tid; j < n_cols; j += stride) {
        p.o2p[j] = SLA_DEV_NONE;
        if (all) p.prices[j] = 0.0;
    }
}

// =============================================================================================================
// Utility kernels
// =============================================================================================================
// init_solve (reference src/solver.rs:218-229): prices = 0, both assignment vectors = NONE.
__global__ void __launch_bounds__(kWideThreads) init_solve_kernel(const Params p, const uint32_t n_rows, const uint32_t n_cols,
                                                                  const int clear_best) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t j = tid; j < n_cols; j += stride) {
        p.prices[j] = 0.0;
        p.o2p[j] = SLA_DEV_NONE;
        if (clear_best) p.best[j] = 0ull;
    }
    for (uint32_t i = tid; i < n_rows; i += stride) p.p2o[i] = SLA_DEV_NONE;
}

// First round of a solve, object side (initialisation of solver.rs:218-229 + the object half of the assignment when the
// first scan ran with all prices zero: it read neither prices nor owners, so they are only written now, behind it): one coalesced sweep over the objects.  An object whose bid word is set takes
// the exact f64 bid of the person the word elected (first round: slot index == person) as its price, records that
// person as its owner and clears the word; every other object gets price 0 and no owner (solver.rs:218-229).  All
// stores are coalesced; the only scattered access is the winner's 8-byte bid, still in the L2 behind the scan.
__global__ void __launch_bounds__(kWideThreads) first_assign_objects_kernel(const Params p) {
    const HotState h = load_hot(p.st);
    if (h.done) return;
    const uint32_t M = h.n_cols;
    const unsigned long long pmask = (1ull << h.pbits) - 1ull;
    const uint32_t base = h.person_base;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint32_t pairs = M / 2u;
    for (uint32_t t = tid; t < pairs; t += stride) {
        const ulonglong2 w = __ldcg(reinterpret_cast<const ulonglong2*>(p.best) + t);
        double2 pr = make_double2(0.0, 0.0);
        uint2 ow = make_uint2(SLA_DEV_NONE, SLA_DEV_NONE);
        if (w.x) { ow.x = (uint32_t)(pmask - (w.x & pmask)) - base; pr.x = __ldcg(p.slot_bid + ow.x); }
        if (w.y) { ow.y = (uint32_t)(pmask - (w.y & pmask)) - base; pr.y = __ldcg(p.slot_bid + ow.y); }
        reinterpret_cast<double2*>(p.prices)[t] = pr;
        reinterpret_cast<uint2*>(p.o2p)[t] = ow;
        if (w.x | w.y) reinterpret_cast<ulonglong2*>(p.best)[t] = make_ulonglong2(0ull, 0ull);
    }
    if (tid == 0 && (M & 1u)) {
        const uint32_t j = M - 1u;
        const unsigned long long w = __ldcg(p.best + j);
        if (w) {
            const uint32_t i = (uint32_t)(pmask - (w & pmask)) - base;
            p.prices[j] = __ldcg(p.slot_bid + i); p.o2p[j] = i; p.best[j] = 0ull;
        } else { p.prices[j] = 0.0; p.o2p[j] = SLA_DEV_NONE; }
    }
}

// Value range + structural validation of an uploaded CSR (ksparse.rs:171-179, symmetric.rs:246, solver.rs:241).
__global__ void __launch_bounds__(kWideThreads) csr_stats_kernel(const uint32_t* __restrict__ row_ptr,
                                                                 const uint32_t* __restrict__ cols,
                                                                 const double* __restrict__ vals, const uint32_t n_rows,
                                                                 const uint32_t n_cols, const unsigned long long nnz,
                                                                 DevCsrStats* out) {
    const unsigned long long tid = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long kmin = ~0ull, kmax = 0ull, bad_c = 0, bad_r = 0, irr = 0, n16 = 0;
    const uint32_t k0 = row_ptr[1] - row_ptr[0];
    // 4 arcs per thread and step: one 128-bit load of column indices, two of values (the arrays are 256 B-aligned)
    const unsigned long long quads = nnz / 4ull;
    for (unsigned long long t = tid; t < quads; t += stride) {
        const uint4 c4 = __ldg(reinterpret_cast<const uint4*>(cols) + t);
        const double2 v01 = __ldg(reinterpret_cast<const double2*>(vals) + 2ull * t);
        const double2 v23 = __ldg(reinterpret_cast<const double2*>(vals) + 2ull * t + 1ull);
        const unsigned long long k0_ = f64_order_key(v01.x), k1_ = f64_order_key(v01.y), k2_ = f64_order_key(v23.x),
                                 k3_ = f64_order_key(v23.y);
        const unsigned long long lo01 = k0_ < k1_ ? k0_ : k1_, lo23 = k2_ < k3_ ? k2_ : k3_;
        const unsigned long long hi01 = k0_ > k1_ ? k0_ : k1_, hi23 = k2_ > k3_ ? k2_ : k3_;
        const unsigned long long lo = lo01 < lo23 ? lo01 : lo23, hi = hi01 > hi23 ? hi01 : hi23;
        kmin = lo < kmin ? lo : kmin;
        kmax = hi > kmax ? hi : kmax;
        bad_c += (c4.x >= n_cols ? 1u : 0u) + (c4.y >= n_cols ? 1u : 0u) + (c4.z >= n_cols ? 1u : 0u) + (c4.w >= n_cols ? 1u : 0u);
        n16 += (is_u16_value(v01.x) ? 0u : 1u) + (is_u16_value(v01.y) ? 0u : 1u) + (is_u16_value(v23.x) ? 0u : 1u) +
               (is_u16_value(v23.y) ? 0u : 1u);
    }
    for (unsigned long long g = quads * 4ull + tid; g < nnz; g += stride) {
        const unsigned long long k = f64_order_key(vals[g]);
        kmin = k < kmin ? k : kmin;
        kmax = k > kmax ? k : kmax;
        bad_c += (cols[g] >= n_cols) ? 1u : 0u;
        n16 += is_u16_value(vals[g]) ? 0u : 1u;
    }
    for (unsigned long long i = tid; i < n_rows; i += stride) {
        const uint32_t a = row_ptr[i], b = row_ptr[i + 1];
        bad_r += (b < a || (unsigned long long)b > nnz) ? 1u : 0u;
        irr += (b - a != k0) ? 1u : 0u;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        const unsigned long long omin = __shfl_xor_sync(0xffffffffu, kmin, m);
        const unsigned long long omax = __shfl_xor_sync(0xffffffffu, kmax, m);
        kmin = omin < kmin ? omin : kmin;
        kmax = omax > kmax ? omax : kmax;
        bad_c += __shfl_xor_sync(0xffffffffu, bad_c, m);
        bad_r += __shfl_xor_sync(0xffffffffu, bad_r, m);
        irr += __shfl_xor_sync(0xffffffffu, irr, m);
        n16 += __shfl_xor_sync(0xffffffffu, n16, m);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&out->min_key, kmin);
        atomicMax(&out->max_key, kmax);
        if (bad_c) atomicAdd(&out->bad_cols, bad_c);
        if (bad_r) atomicAdd(&out->bad_rows, bad_r);
        if (irr) atomicAdd(&out->irregular_rows, irr);
        if (n16) atomicAdd(&out->not_u16, n16);
    }
}

// u16 mirror of values that are all integers in [0, 65535] (csr_stats_kernel: not_u16 == 0) for a CSR that did not
// cross PCIe narrow (generated in HBM, handed over in device memory): the uniform-degree scans then read 6 instead of
// 12 bytes per arc, exactly as behind a u16 upload.  8 values per thread and pass.
__global__ void __launch_bounds__(kWideThreads) narrow_mirror_kernel(const double* __restrict__ src, uint16_t* __restrict__ dst,
                                                                     const size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 8u;
    for (size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 8u; g < n; g += stride) {
        if (g + 8u <= n) {
            uint32_t w[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const double2 d = *reinterpret_cast<const double2*>(src + g + 2 * u);
                w[u] = (uint32_t)d.x | ((uint32_t)d.y << 16);
            }
            *reinterpret_cast<uint4*>(dst + g) = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (size_t u = g; u < n; ++u) dst[u] = (uint16_t)(uint32_t)src[u];
        }
    }
}

// get_objective on the resident solution (reference src/solver.rs:110-142): per-block partial sums of the
// effective values of the chosen arcs; the host adds the partials in block order (deterministic).
__global__ void __launch_bounds__(kWideThreads) objective_kernel(const Params p, const uint32_t n_rows, const uint32_t sign_flip,
                                                                 double* __restrict__ partial) {
    double acc = 0.0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t i = tid; i < n_rows; i += stride) {
        const uint32_t j = p.p2o[i];
        if (j == SLA_DEV_NONE) continue;
        const uint32_t a = p.row_ptr[i], b = p.row_ptr[i + 1];
        for (uint32_t g = a; g < b; ++g) {
            if (p.cols[g] == j) {
                const double raw = p.vals[g];
                acc += __hiloint2double(__double2hiint(raw) ^ (int)sign_flip, __double2loint(raw));
            }
        }
    }
    __shared__ double s_part[kWideThreads / 32];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, m);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < kWideThreads / 32; ++w) t += s_part[w];
        partial[blockIdx.x] = t;
    }
}

// Stand-alone eps-CS check on the resident solution (reference src/solver.rs:154-189); out[0] |= violated.
__global__ void __launch_bounds__(kWideThreads) ecs_check_kernel(const Params p, const uint32_t n_rows, const uint32_t n_cols,
                                                                 const uint32_t sign_flip, const double eps, const double tol,
                                                                 uint32_t* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    bool violated = false;
    for (uint32_t i = tid; i < n_rows; i += stride) {
        const uint32_t j = p.p2o[i];
        if (j >= n_cols) { violated = true; continue; }
        const uint32_t a = p.row_ptr[i], b = p.row_ptr[i + 1];
        double chosen = neg_inf();
        for (uint32_t g = a; g < b; ++g)
            if (p.cols[g] == j) {
                const double raw = p.vals[g];
                chosen = __hiloint2double(__double2hiint(raw) ^ (int)sign_flip, __double2loint(raw));
            }
        const double lhs = chosen - p.prices[j] + tol;
        for (uint32_t g = a; g < b; ++g) {
            const double raw = p.vals[g];
            const double v = __hiloint2double(__double2hiint(raw) ^ (int)sign_flip, __double2loint(raw));
            if (lhs < v - p.prices[p.cols[g]] - eps) violated = true;
        }
    }
    if (violated) out[0] = 1u;
}

// Matching validator: out[0] = unassigned persons, out[1] = inconsistencies between the two vectors.
__global__ void __launch_bounds__(kWideThreads) validate_kernel(const Params p, const uint32_t n_rows, const uint32_t n_cols,
                                                                uint32_t* out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    uint32_t unassigned = 0, bad = 0;
    for (uint32_t i = tid; i < n_rows; i += stride) {
        const uint32_t j = p.p2o[i];
        if (j == SLA_DEV_NONE) unassigned += 1;
        else if (j >= n_cols || p.o2p[j] != i) bad += 1;
    }
    for (uint32_t j = tid; j < n_cols; j += stride) {
        const uint32_t i = p.o2p[j];
        if (i != SLA_DEV_NONE && (i >= n_rows || p.p2o[i] != j)) bad += 1;
    }
    if (unassigned) atomicAdd(out + 0, unassigned);
    if (bad) atomicAdd(out + 1, bad);
}

// Synthetic instance generator (synth.h), one thread per row.
__global__ void __launch_bounds__(kWideThreads) generate_kernel(const sla_synth::Spec spec, const uint32_t row_begin,
                                                                const uint32_t row_count, uint32_t* __restrict__ row_ptr,
                                                                uint32_t* __restrict__ cols, double* __restrict__ vals) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    for (uint32_t r = tid; r < row_count; r += stride) {
        const size_t off = (size_t)r * spec.k;
        sla_synth::make_row(spec, row_begin + r, cols + off, vals + off);
        row_ptr[r] = (uint32_t)off;
    }
    if (tid == 0) row_ptr[row_count] = row_count * spec.k;
}

}  // namespace sla
