/*
 * jacobi_model.c -- TEST INFRASTRUCTURE ONLY (lives beside the CPU oracle, never on the product path).
 *
 * A sequential CPU model of the *device* algorithm (DESIGN.md "Round semantics"): synchronous Jacobi auction
 * rounds in which every queued person bids against frozen prices, conflicts are resolved per object by the
 * maximum of a packed 64-bit (bid key, person) word, the winner installs its own exact f64 bid as the price,
 * evicted owners and losers form the next queue.  It is NOT a restatement of the reference (that is
 * sla_oracle.c); it exists so that the CUDA kernels can be checked bit for bit (prices, assignment vectors,
 * round/bid/arc counters), independent of thread scheduling, and so that the multi-rank host logic can be
 * exercised on CPU (gloo) with this model standing in for the kernels.
 *
 * The arithmetic of one bid is the reference's: src/ksparse.rs:199-227 and src/symmetric.rs:361-378.
 *
 * Khosla on square instances: the device runs the rounds under an eps-schedule (c/2, x0.15, ..., the caller's eps) and
 * falls back to the plain rounds from scratch as soon as a phase drops anybody at the price threshold; this model does
 * the same (jm_set_khosla_scaling(0) gives the plain rounds throughout, like the library option "khosla_scaling" = 0).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define JM_NONE 0xFFFFFFFFu
#define JM_KHOSLA 0
#define JM_FORWARD 1

typedef struct {
    uint32_t num_unassigned, nits, nreductions, optimal_soln_found;
    double eps;
    uint64_t rounds, bids, bid_arcs;
    uint32_t dropped, values_negated;
    uint32_t restarts, reserved_;   /* restarts: 1 when the Khosla eps-schedule was abandoned for the plain rounds */
} jm_stats;

/* 1: Khosla rounds on square instances run under an eps-schedule (set by the tests / the bindings) */
int jm_khosla_scaling = 1;
void jm_set_khosla_scaling(int on) { jm_khosla_scaling = on; }

static uint32_t person_bits(uint32_t num_rows) {
    uint32_t m = num_rows > 1 ? num_rows - 1 : 1;
    return 32u - (uint32_t)__builtin_clz(m);
}

/* order-preserving map of an f64 onto u64, truncated so that the person id fits below it; bit 63 stays 0 */
uint64_t jm_pack_bid(double bid, uint32_t person, uint32_t pbits) {
    uint64_t u;
    memcpy(&u, &bid, 8);
    u ^= (u >> 63) ? ~(uint64_t)0 : ((uint64_t)1 << 63);
    uint64_t pmask = ((uint64_t)1 << pbits) - 1;
    return ((u >> (pbits + 1)) << pbits) | (pmask - person);
}

typedef struct {
    uint32_t jbest;
    double best_profit, best_value, second_profit;
} scan_t;

/* choice rule: strict '>' so the lowest row position wins ties (ksparse.rs:206, symmetric.rs:367) */
static scan_t scan_row(const uint32_t *cols, const double *vals, size_t a, size_t b, const double *prices,
                       double sign) {
    scan_t r = {0u, -INFINITY, -INFINITY, -INFINITY};
    for (size_t g = a; g < b; ++g) {
        double v = sign < 0 ? -vals[g] : vals[g];
        double profit = v - prices[cols[g]];
        if (profit > r.best_profit) {
            r.jbest = cols[g];
            r.second_profit = r.best_profit;
            r.best_profit = profit;
            r.best_value = v;
        } else if (profit > r.second_profit) {
            r.second_profit = profit;
        }
    }
    return r;
}

static int ecs_ok(uint32_t n_rows, const uint32_t *row_ptr, const uint32_t *cols, const double *vals, double sign,
                  const double *prices, const uint32_t *p2o, double eps, double tol) {
    /* solver.rs:154-189 */
    for (uint32_t i = 0; i < n_rows; ++i) {
        uint32_t j = p2o[i];
        double chosen = -INFINITY;
        for (size_t g = row_ptr[i]; g < row_ptr[i + 1]; ++g)
            if (cols[g] == j) chosen = sign < 0 ? -vals[g] : vals[g];
        double lhs = chosen - prices[j] + tol;
        for (size_t g = row_ptr[i]; g < row_ptr[i + 1]; ++g) {
            double v = sign < 0 ? -vals[g] : vals[g];
            if (lhs < v - prices[cols[g]] - eps) return 0;
        }
    }
    return 1;
}

static double toleration(double c) {
    double l = log2(c + 1e-7);
    uint32_t li = !(l > 0.0) ? 0u : (l >= 4294967295.0 ? 4294967295u : (uint32_t)l);
    uint32_t e = 53u - li;
    uint64_t p = e < 64 ? ((uint64_t)1 << e) : 0;
    return 1.0 / (double)p;
}

/*
 * One synchronous round over queue[0..qlen).  Returns the next queue length (written to next_queue).
 * slot_obj / slot_bid are scratch of qlen entries; best[] must be all-zero on entry and is all-zero on exit.
 */
static uint32_t jacobi_round(int algo, const uint32_t *row_ptr, const uint32_t *cols, const double *vals, double sign,
                             double eps, double threshold, uint32_t pbits, double *prices, uint32_t *p2o,
                             uint32_t *o2p, uint64_t *best, const uint32_t *queue, uint32_t qlen,
                             uint32_t *next_queue, uint32_t *slot_obj, double *slot_bid, jm_stats *st) {
    for (uint32_t q = 0; q < qlen; ++q) {
        uint32_t i = queue[q];
        scan_t r = scan_row(cols, vals, row_ptr[i], row_ptr[i + 1], prices, sign);
        st->bid_arcs += row_ptr[i + 1] - row_ptr[i];
        st->bids += 1;
        double bid;
        if (algo == JM_KHOSLA) {
            if (prices[r.jbest] > threshold) { /* ksparse.rs:218-220 */
                slot_obj[q] = JM_NONE;
                slot_bid[q] = 0.0;
                st->dropped += 1;
                continue;
            }
            bid = isfinite(r.second_profit) ? r.best_value - r.second_profit + eps : prices[r.jbest] + eps;
        } else {
            bid = r.best_value - r.second_profit + eps; /* symmetric.rs:378 */
        }
        slot_obj[q] = r.jbest;
        slot_bid[q] = bid;
        if (bid != bid) continue; /* NaN never wins (symmetric.rs:394) */
        uint64_t w = jm_pack_bid(bid, i, pbits);
        if (w > best[r.jbest]) best[r.jbest] = w;
    }
    uint32_t nq = 0;
    /* two passes so that the reset of best[] cannot disturb a later reader, mirroring the kernel's word test */
    for (uint32_t q = 0; q < qlen; ++q) {
        uint32_t i = queue[q], j = slot_obj[q];
        if (j == JM_NONE) continue; /* dropped */
        double bid = slot_bid[q];
        int won = (bid == bid) && best[j] == jm_pack_bid(bid, i, pbits);
        if (won) {
            uint32_t prev = o2p[j];
            prices[j] = bid;
            o2p[j] = i;
            p2o[i] = j;
            if (prev != JM_NONE) {
                p2o[prev] = JM_NONE;
                next_queue[nq++] = prev;
            }
        } else {
            next_queue[nq++] = i;
        }
    }
    for (uint32_t q = 0; q < qlen; ++q)
        if (slot_obj[q] != JM_NONE) best[slot_obj[q]] = 0;
    st->rounds += 1;
    return nq;
}

/*
 * Full solve.  algo: 0 Khosla, 1 Forward.  eps / start_eps NaN => default; max_iterations 0 => 100000.
 * Outputs: p2o[N], o2p[M], prices[M].  vals are NOT modified (sign applied on the fly); st->values_negated says
 * whether the reference would have negated them in place (solver.rs:209-216).
 */
int jm_solve(int algo, uint32_t n_rows, uint32_t n_cols, const uint32_t *row_ptr, const uint32_t *cols,
             const double *vals, int maximize, double eps_in, double start_eps_in, uint32_t max_iterations,
             uint32_t *p2o, uint32_t *o2p, double *prices, jm_stats *st) {
    size_t nnz = row_ptr[n_rows];
    memset(st, 0, sizeof *st);
    int positive = (nnz ? vals[0] : 0.0) >= 0.0;
    int negate = (maximize != 0) ^ positive;
    double sign = negate ? -1.0 : 1.0;
    st->values_negated = (uint32_t)negate;

    for (uint32_t j = 0; j < n_cols; ++j) { prices[j] = 0.0; o2p[j] = JM_NONE; }
    for (uint32_t i = 0; i < n_rows; ++i) p2o[i] = JM_NONE;

    double w_min = INFINITY, w_max = -INFINITY, c = 0.0;
    for (size_t a = 0; a < nnz; ++a) {
        double v = negate ? -vals[a] : vals[a];
        w_min = w_min < v ? w_min : v;
        w_max = w_max > v ? w_max : v;
        c = fmax(c, fabs(v));
    }

    uint32_t pbits = person_bits(n_rows);
    uint64_t *best = (uint64_t *)calloc(n_cols, sizeof(uint64_t));
    uint32_t *qa = (uint32_t *)malloc((size_t)n_rows * 4), *qb = (uint32_t *)malloc((size_t)n_rows * 4);
    uint32_t *slot_obj = (uint32_t *)malloc((size_t)n_rows * 4);
    double *slot_bid = (double *)malloc((size_t)n_rows * 8);
    for (uint32_t i = 0; i < n_rows; ++i) qa[i] = i;
    uint32_t qlen = n_rows;

    if (algo == JM_KHOSLA) {
        double m = (double)n_cols;
        double target = isnan(eps_in) ? 1.0 / m : eps_in;
        st->eps = target;
        double threshold = (m / 2.0) * (w_max - w_min + target);
        /* eps-schedule of the device's Khosla rounds on square instances (DESIGN.md "Khosla under an eps-schedule"):
         * phases at c/2, x0.15, ... each restarting from the kept prices, the last one at exactly the caller's eps */
        double eps = target;
        int scaled = jm_khosla_scaling && n_rows == n_cols && c / 2.0 > target;
        if (scaled) eps = c / 2.0;
        for (;;) {
            while (qlen > 0) {
                qlen = jacobi_round(algo, row_ptr, cols, vals, sign, eps, threshold, pbits, prices, p2o, o2p, best, qa,
                                    qlen, qb, slot_obj, slot_bid, st);
                uint32_t *t = qa; qa = qb; qb = t;
            }
            if (!scaled) break;
            if (st->dropped != 0) {
                /* somebody hit the price threshold under the schedule: start over with the plain rounds, which alone
                 * define what the threshold does on an instance without a perfect matching */
                scaled = 0;
                st->restarts = 1;
                eps = target;
                st->dropped = 0;
                for (uint32_t j = 0; j < n_cols; ++j) prices[j] = 0.0;
            } else if (eps > target) {
                eps *= 0.15;
                if (eps < target) eps = target;
                st->nreductions += 1;
            } else {
                break;
            }
            for (uint32_t i = 0; i < n_rows; ++i) { p2o[i] = JM_NONE; qa[i] = i; }
            for (uint32_t j = 0; j < n_cols; ++j) o2p[j] = JM_NONE;
            qlen = n_rows;
        }
        st->nits = (uint32_t)st->bids;
        st->num_unassigned = st->dropped;
    } else {
        double target = isnan(eps_in) ? 1.0 / (double)n_rows : eps_in;
        uint32_t max_it = max_iterations ? max_iterations : 100000u;
        double tol = toleration(c);
        int start_opt = !isnan(start_eps_in) ? (start_eps_in < target) : 0;
        double eps;
        if (n_rows != n_cols) { start_opt = 1; eps = target - 2.220446049250313e-16; }
        else eps = !isnan(start_eps_in) ? start_eps_in : c / 2.0;
        for (;;) {
            qlen = jacobi_round(algo, row_ptr, cols, vals, sign, eps, 0.0, pbits, prices, p2o, o2p, best, qa, qlen,
                                qb, slot_obj, slot_bid, st);
            uint32_t *t = qa; qa = qb; qb = t;
            st->nits += 1;
            if (qlen == 0) {
                int optimal = start_opt || ecs_ok(n_rows, row_ptr, cols, vals, sign, prices, p2o, target, tol);
                if (optimal) { st->optimal_soln_found = 1; break; }
                if (eps < target) break;
                eps *= 0.15;
                for (uint32_t i = 0; i < n_rows; ++i) { p2o[i] = JM_NONE; qa[i] = i; }
                for (uint32_t j = 0; j < n_cols; ++j) o2p[j] = JM_NONE;
                qlen = n_rows;
                st->nreductions += 1;
            }
            if (st->nits >= max_it) break;
        }
        st->eps = eps;
        st->num_unassigned = qlen;
    }
    free(best); free(qa); free(qb); free(slot_obj); free(slot_bid);
    return 0;
}
