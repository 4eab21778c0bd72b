/*
 * sla_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY, never shipped, never on the product path).
 *
 * A plain-C restatement of the sequential CPU algorithms of DXist/sparse_linear_assignment v0.1.5,
 * used as the checker for the CUDA path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED.  The Rust reference cannot be compiled in this image (no cargo/rustc), so this
 * is a "port" oracle; it is pinned by reproducing every golden value the reference's own tests hold
 * (tests/test_oracle_goldens.py): src/solver.rs:296, 332-336, 348-388, 435; doctest src/ksparse.rs:28-39;
 * src/symmetric.rs:516-534.
 *
 * Each function cites the reference lines it follows.  Index type I (u16/u32 in the reference,
 * src/solution.rs:16-17) is carried as uint32_t plus `imax` (= I::MAX) so that the overflow /
 * sentinel behaviour of either width can be reproduced.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK 0
#define ORC_ERR 1

typedef struct { uint32_t *p; size_t len, cap; } vec_u32;
typedef struct { double *p; size_t len, cap; } vec_f64;

static void vu_reserve(vec_u32 *v, size_t n) {
    if (n <= v->cap) return;
    size_t c = v->cap ? v->cap : 16;
    while (c < n) c *= 2;
    v->p = (uint32_t *)realloc(v->p, c * sizeof(uint32_t));
    v->cap = c;
}
static void vf_reserve(vec_f64 *v, size_t n) {
    if (n <= v->cap) return;
    size_t c = v->cap ? v->cap : 16;
    while (c < n) c *= 2;
    v->p = (double *)realloc(v->p, c * sizeof(double));
    v->cap = c;
}
static void vu_push(vec_u32 *v, uint32_t x) { vu_reserve(v, v->len + 1); v->p[v->len++] = x; }
static void vf_push(vec_f64 *v, double x) { vf_reserve(v, v->len + 1); v->p[v->len++] = x; }
static void vu_fill(vec_u32 *v, size_t n, uint32_t x) {
    vu_reserve(v, n); v->len = n;
    for (size_t i = 0; i < n; ++i) v->p[i] = x;
}
static void vf_fill(vec_f64 *v, size_t n, double x) {
    vf_reserve(v, n); v->len = n;
    for (size_t i = 0; i < n; ++i) v->p[i] = x;
}

/* AuctionSolution<I>  (src/solution.rs:22-53) */
typedef struct {
    vec_u32 person_to_object, object_to_person;
    uint32_t num_unassigned;
    double eps;
} orc_solution;

/* Union of the KhoslaSolver (src/ksparse.rs:73-85) and ForwardAuctionSolver (src/symmetric.rs:75-98) state. */
typedef struct {
    uint32_t imax; /* I::MAX */
    uint32_t num_rows, num_cols;
    vec_f64 prices, values;
    vec_u32 i_starts_stops, j_counts, column_indices;
    /* khosla */
    vec_u32 ustack;
    /* forward */
    uint32_t max_iterations;
    vec_f64 best_bids;
    vec_u32 best_bidders, unassigned_people, person_to_assignment_idx;
    /* public counters */
    uint32_t nits, nreductions;
    int optimal_soln_found;
    /* instrumentation that the reference does not have: arcs examined by the bid scans */
    uint64_t bid_arcs;
} orc_solver;

orc_solver *orc_solver_new(uint32_t imax, size_t row_cap, size_t col_cap, size_t arc_cap) {
    /* ksparse.rs:88-107 / symmetric.rs:101-130: capacities only */
    orc_solver *s = (orc_solver *)calloc(1, sizeof(orc_solver));
    s->imax = imax;
    vu_reserve(&s->i_starts_stops, row_cap + 1);
    vu_reserve(&s->j_counts, row_cap);
    vf_reserve(&s->prices, col_cap);
    vu_reserve(&s->column_indices, arc_cap);
    vf_reserve(&s->values, arc_cap);
    s->max_iterations = 100000; /* symmetric.rs:190 */
    return s;
}
void orc_solver_free(orc_solver *s) {
    if (!s) return;
    free(s->prices.p); free(s->values.p); free(s->i_starts_stops.p); free(s->j_counts.p);
    free(s->column_indices.p); free(s->ustack.p); free(s->best_bids.p); free(s->best_bidders.p);
    free(s->unassigned_people.p); free(s->person_to_assignment_idx.p);
    free(s);
}
orc_solution *orc_solution_new(uint32_t imax) {
    /* solution.rs:46-53 */
    orc_solution *z = (orc_solution *)calloc(1, sizeof(orc_solution));
    z->eps = NAN;
    z->num_unassigned = imax;
    return z;
}
void orc_solution_free(orc_solution *z) {
    if (!z) return;
    free(z->person_to_object.p); free(z->object_to_person.p); free(z);
}

/* solver.rs:191-205 */
int orc_init(orc_solver *s, uint32_t num_rows, uint32_t num_cols) {
    if (!(num_rows <= num_cols)) return ORC_ERR;
    if (!(num_rows < s->imax)) return ORC_ERR;
    s->num_rows = num_rows;
    s->num_cols = num_cols;
    s->i_starts_stops.len = 0;
    vu_push(&s->i_starts_stops, 0);
    vu_push(&s->i_starts_stops, 0);
    s->j_counts.len = 0;
    vu_push(&s->j_counts, 0);
    s->column_indices.len = 0;
    s->values.len = 0;
    return ORC_OK;
}

/* solver.rs:41-66 */
int orc_add_value(orc_solver *s, uint32_t row, uint32_t column, double value) {
    size_t current_row = s->j_counts.len - 1;
    size_t r = row;
    if (!(r == current_row || r == current_row + 1)) return ORC_ERR;
    uint32_t prev = s->i_starts_stops.p[current_row + 1];
    if (prev >= s->imax) return ORC_ERR; /* checked_add(1) overflow */
    uint32_t cumulative = prev + 1;
    if (r > current_row) {
        if (!(s->j_counts.p[current_row] > 0)) return ORC_ERR;
        vu_push(&s->i_starts_stops, cumulative);
        vu_push(&s->j_counts, 1);
    } else {
        s->i_starts_stops.p[current_row + 1] = cumulative;
        s->j_counts.p[current_row] += 1;
    }
    vu_push(&s->column_indices, column);
    vf_push(&s->values, value);
    return ORC_OK;
}

/* solver.rs:69-101 */
int orc_extend_from_values(orc_solver *s, uint32_t row, const uint32_t *columns, size_t ncolumns,
                           const double *values, size_t nvalues) {
    if (ncolumns != nvalues) return ORC_ERR;
    size_t current_row = s->j_counts.len - 1;
    size_t r = row;
    if (!(r == current_row || r == current_row + 1)) return ORC_ERR;
    if (ncolumns > (size_t)s->imax) return ORC_ERR; /* I::from_usize */
    uint32_t inc = (uint32_t)ncolumns;
    uint32_t prev = s->i_starts_stops.p[current_row + 1];
    if ((uint64_t)prev + inc > (uint64_t)s->imax) return ORC_ERR; /* checked_add */
    uint32_t cumulative = prev + inc;
    if (r > current_row) {
        if (!(s->j_counts.p[current_row] > 0)) return ORC_ERR;
        vu_push(&s->i_starts_stops, cumulative);
        vu_push(&s->j_counts, inc);
    } else {
        s->i_starts_stops.p[current_row + 1] = cumulative;
        s->j_counts.p[current_row] += inc;
    }
    vu_reserve(&s->column_indices, s->column_indices.len + ncolumns);
    memcpy(s->column_indices.p + s->column_indices.len, columns, ncolumns * sizeof(uint32_t));
    s->column_indices.len += ncolumns;
    vf_reserve(&s->values, s->values.len + nvalues);
    memcpy(s->values.p + s->values.len, values, nvalues * sizeof(double));
    s->values.len += nvalues;
    return ORC_OK;
}

/* Convenience for large inputs: init + one extend_from_values per row of a CSR (same calls a user makes). */
int orc_load_csr(orc_solver *s, uint32_t num_rows, uint32_t num_cols, const uint32_t *row_ptr,
                 const uint32_t *cols, const double *vals) {
    int rc = orc_init(s, num_rows, num_cols);
    if (rc) return rc;
    for (uint32_t i = 0; i < num_rows; ++i) {
        size_t a = row_ptr[i], b = row_ptr[i + 1];
        rc = orc_extend_from_values(s, i, cols + a, b - a, vals + a, b - a);
        if (rc) return rc;
    }
    return ORC_OK;
}

size_t orc_num_of_arcs(const orc_solver *s) { return s->column_indices.len; } /* solver.rs:104-106 */

/* solver.rs:232-243 (the debug_assert on column bounds is compiled out in release) */
int orc_validate_input(const orc_solver *s) {
    size_t arcs = s->column_indices.len;
    if (!(arcs > 0)) return ORC_ERR;
    if (!(s->num_rows > 0 && s->num_cols > 0)) return ORC_ERR;
    if (!(arcs < (size_t)s->imax)) return ORC_ERR;
    if (!(arcs == s->column_indices.len && s->column_indices.len == s->values.len)) return ORC_ERR;
    return ORC_OK;
}

/* solver.rs:207-230 */
static void trait_init_solve(orc_solver *s, orc_solution *z, int maximize) {
    double first = s->values.len ? s->values.p[0] : 0.0;
    int positive_values = first >= 0.0;
    if ((maximize != 0) ^ positive_values) {
        for (size_t a = 0; a < s->values.len; ++a) s->values.p[a] *= -1.0;
    }
    vf_fill(&s->prices, s->num_cols, 0.0);
    vu_fill(&z->person_to_object, s->num_rows, s->imax);
    vu_fill(&z->object_to_person, s->num_cols, s->imax);
    z->num_unassigned = s->num_rows;
}

/* solver.rs:110-142 */
double orc_get_objective(const orc_solver *s, const orc_solution *z) {
    double first = s->values.len ? s->values.p[0] : 0.0;
    int positive_values = first >= 0.0;
    double obj = 0.0;
    for (uint32_t i = 0; i < s->num_rows; ++i) {
        uint32_t j = z->person_to_object.p[i];
        if (j == s->imax) continue;
        uint32_t n = s->j_counts.p[i];
        uint32_t start = s->i_starts_stops.p[i];
        for (uint32_t idx = 0; idx < n; ++idx) {
            size_t g = (size_t)start + idx;
            if (s->column_indices.p[g] == j) {
                if (positive_values) obj += s->values.p[g];
                else obj -= s->values.p[g];
            }
        }
    }
    return obj;
}

/* solver.rs:144-146: 1 / 2^(53 - (log2(c + 1e-7) as u32)); `as u32` saturates (negative / NaN -> 0) */
double orc_get_toleration(double max_abs_cost) {
    double l = log2(max_abs_cost + 1e-7);
    uint32_t li;
    if (!(l > 0.0)) li = 0;
    else if (l >= 4294967295.0) li = 4294967295u;
    else li = (uint32_t)l;
    uint32_t e = 53u - li; /* wraps like release-mode Rust when li > 53 */
    uint64_t p = (e < 64) ? ((uint64_t)1 << e) : 0; /* 2_u64.pow(e) overflows for e >= 64 */
    return 1.0 / (double)p;
}

/* solver.rs:154-189 */
int orc_ecs_satisfied(const orc_solver *s, const uint32_t *person_to_object, double eps, double toleration) {
    for (uint32_t i = 0; i < s->num_rows; ++i) {
        uint32_t n = s->j_counts.p[i];
        uint32_t start = s->i_starts_stops.p[i];
        uint32_t j = person_to_object[i];
        double chosen_value = -INFINITY;
        for (uint32_t idx = 0; idx < n; ++idx) {
            size_t g = (size_t)start + idx;
            if (s->column_indices.p[g] == j) chosen_value = s->values.p[g];
        }
        double lhs = chosen_value - s->prices.p[j] + toleration;
        for (uint32_t idx = 0; idx < n; ++idx) {
            size_t g = (size_t)start + idx;
            size_t k = s->column_indices.p[g];
            double value = s->values.p[g];
            if (lhs < value - s->prices.p[k] - eps) return 0;
        }
    }
    return 1;
}

/* KhoslaSolver::solve  (ksparse.rs:153-251; inherent init_solve 253-260).  eps NaN => None. */
int orc_khosla_solve(orc_solver *s, orc_solution *z, int maximize, double eps_or_nan) {
    if (orc_validate_input(s)) return ORC_ERR;
    trait_init_solve(s, z, maximize);
    /* ustack = [n-1, ..., 1, 0]  so that pop() yields 0, 1, 2, ... */
    s->ustack.len = 0;
    for (uint32_t i = s->num_rows; i-- > 0;) vu_push(&s->ustack, i);

    double num_cols_f = (double)s->num_cols;
    double eps = isnan(eps_or_nan) ? 1.0 / num_cols_f : eps_or_nan;
    z->eps = eps;

    double w_min = INFINITY, w_max = -INFINITY;
    for (size_t a = 0; a < s->values.len; ++a) {
        double el = s->values.p[a];
        w_min = (w_min < el) ? w_min : el;
        w_max = (w_max > el) ? w_max : el;
    }
    double price_threshold = (num_cols_f / 2.0) * (w_max - w_min + eps);

    s->nits = 0;
    s->bid_arcs = 0;
    const uint32_t *col = s->column_indices.p;
    const double *val = s->values.p;
    double *prices = s->prices.p;
    while (s->ustack.len > 0) {
        uint32_t u = s->ustack.p[--s->ustack.len];
        s->nits += 1;
        size_t start = s->i_starts_stops.p[u];
        size_t n = s->j_counts.p[u];
        double max_profit = -INFINITY, max_edge_value = -INFINITY, second_max_profit = -INFINITY;
        uint32_t matched_v = 0;
        for (size_t idx = 0; idx < n; ++idx) {
            size_t g = start + idx;
            uint32_t j = col[g];
            double edge_value = val[g];
            double profit = edge_value - prices[j];
            if (profit > max_profit) {
                matched_v = j;
                second_max_profit = max_profit;
                max_profit = profit;
                max_edge_value = edge_value;
            } else if (profit > second_max_profit) {
                second_max_profit = profit;
            }
        }
        s->bid_arcs += n;
        if (prices[matched_v] > price_threshold) continue; /* dropped for good: ksparse.rs:218-220 */
        if (isfinite(second_max_profit)) prices[matched_v] = max_edge_value - second_max_profit + eps;
        else prices[matched_v] += eps;

        uint32_t moved_out = z->object_to_person.p[matched_v];
        if (moved_out != s->imax) {
            z->person_to_object.p[moved_out] = s->imax;
            z->num_unassigned += 1;
            vu_push(&s->ustack, moved_out);
        }
        z->person_to_object.p[u] = matched_v;
        z->object_to_person.p[matched_v] = u;
        z->num_unassigned -= 1;
    }
    return ORC_OK;
}

/* push_all_left  (symmetric.rs:471-508).  Reads data[right_track] before the bound test, like the original;
 * `data_len` lets the C version stay memory-safe where the Rust one would panic. */
void orc_push_all_left(uint32_t *data, size_t data_len, uint32_t *mapper, uint32_t num_ints, uint32_t size,
                       uint32_t imax) {
    if (num_ints == 0) return;
    uint32_t left = 0, right = num_ints;
    while (left < num_ints) {
        if (data[left] == imax) {
            while (right < data_len && data[right] == imax && right < size) right += 1;
            if (right >= data_len) return; /* Rust: index out of bounds panic */
            uint32_t i = data[right];
            data[left] = i;
            data[right] = imax;
            mapper[i] = left;
        }
        left += 1;
    }
}

/* ForwardAuctionSolver::bid_and_assign  (symmetric.rs:334-468) */
static void forward_bid_and_assign(orc_solver *s, orc_solution *z, vec_u32 *bidders, vec_u32 *objects_bidded,
                                   vec_f64 *bids) {
    size_t num_bidders = z->num_unassigned;
    vu_fill(bidders, num_bidders, s->imax);
    vu_fill(objects_bidded, num_bidders, s->imax);
    vf_fill(bids, num_bidders, -INFINITY);
    const uint32_t *col = s->column_indices.p;
    const double *val = s->values.p;
    double *prices = s->prices.p;

    /* bidding phase: 343-384 */
    for (size_t nb = 0; nb < num_bidders; ++nb) {
        uint32_t i = s->unassigned_people.p[nb];
        size_t n = s->j_counts.p[i];
        size_t start = s->i_starts_stops.p[i];
        uint32_t jbest = 0;
        double max_edge_value = -INFINITY, max_profit = -INFINITY, second_max_profit = -INFINITY;
        for (size_t idx = 0; idx < n; ++idx) {
            size_t g = start + idx;
            uint32_t j = col[g];
            double edge_value = val[g];
            double profit = edge_value - prices[j];
            if (profit > max_profit) {
                jbest = j;
                second_max_profit = max_profit;
                max_profit = profit;
                max_edge_value = edge_value;
            } else if (profit > second_max_profit) {
                second_max_profit = profit;
            }
        }
        s->bid_arcs += n;
        double bbest = max_edge_value - second_max_profit + z->eps;
        bidders->p[nb] = i;
        bids->p[nb] = bbest;
        objects_bidded->p[nb] = jbest;
    }

    /* conflict resolution: 386-405 */
    size_t num_successful_bids = 0;
    for (size_t n = 0; n < num_bidders; ++n) {
        uint32_t i = bidders->p[n];
        double bid_val = bids->p[n];
        size_t jbid = objects_bidded->p[n];
        if (bid_val > s->best_bids.p[jbid]) {
            if (s->best_bidders.p[jbid] == s->imax) num_successful_bids += 1;
            s->best_bids.p[jbid] = bid_val;
            s->best_bidders.p[jbid] = i;
        }
    }

    /* assignment phase: 409-457 */
    uint32_t to_unassign = 0, to_assign = 0;
    size_t bid_ctr = 0;
    for (uint32_t j = 0; j < s->num_cols; ++j) {
        uint32_t i = s->best_bidders.p[j];
        if (i != s->imax) {
            prices[j] = s->best_bids.p[j];
            uint32_t aidx = s->person_to_assignment_idx.p[i];
            uint32_t prev_i = z->object_to_person.p[j];
            if (prev_i != s->imax) {
                to_unassign += 1;
                z->person_to_object.p[prev_i] = s->imax;
                s->person_to_assignment_idx.p[i] = s->imax;
                s->person_to_assignment_idx.p[prev_i] = aidx;
                s->unassigned_people.p[aidx] = prev_i;
            } else {
                s->unassigned_people.p[aidx] = s->imax;
                s->person_to_assignment_idx.p[i] = s->imax;
            }
            to_assign += 1;
            z->person_to_object.p[i] = j;
            z->object_to_person.p[j] = i;
            s->best_bidders.p[j] = s->imax;
            s->best_bids.p[j] = -INFINITY;
            bid_ctr += 1;
            if (bid_ctr >= num_successful_bids) break;
        }
    }
    z->num_unassigned += to_unassign;
    z->num_unassigned -= to_assign;
    orc_push_all_left(s->unassigned_people.p, s->unassigned_people.len, s->person_to_assignment_idx.p,
                      z->num_unassigned, s->num_cols, s->imax);
}

/* ForwardAuctionSolver::solve_with_params (symmetric.rs:217-332; inherent init_solve 192-215).
 * NaN => None for eps / start_eps; max_iterations 0 => None. */
int orc_forward_solve(orc_solver *s, orc_solution *z, int maximize, double eps_or_nan, double start_eps_or_nan,
                      uint32_t max_iterations_or_0) {
    if (orc_validate_input(s)) return ORC_ERR;
    trait_init_solve(s, z, maximize);
    s->nits = 0;
    s->nreductions = 0;
    s->optimal_soln_found = 0;
    s->bid_arcs = 0;
    vf_fill(&s->best_bids, s->num_cols, -INFINITY);
    vu_fill(&s->best_bidders, s->num_cols, s->imax);
    vu_fill(&s->unassigned_people, s->num_rows, 0);
    vu_fill(&s->person_to_assignment_idx, s->num_rows, 0);
    for (uint32_t i = 0; i < s->num_rows; ++i) {
        s->unassigned_people.p[i] = i;
        s->person_to_assignment_idx.p[i] = i;
    }

    double float_num_rows = (double)s->num_rows;
    double target_eps = isnan(eps_or_nan) ? 1.0 / float_num_rows : eps_or_nan;
    s->max_iterations = max_iterations_or_0 ? max_iterations_or_0 : 100000u;

    double c = 0.0;
    for (size_t a = 0; a < s->values.len; ++a) c = fmax(c, fabs(s->values.p[a]));
    double toleration = orc_get_toleration(c);

    int start_from_optimal_eps = !isnan(start_eps_or_nan) ? (start_eps_or_nan < target_eps) : 0;
    if (s->num_rows != s->num_cols) {
        start_from_optimal_eps = 1;
        z->eps = target_eps - 2.220446049250313e-16; /* f64::EPSILON */
    } else {
        z->eps = !isnan(start_eps_or_nan) ? start_eps_or_nan : c / 2.0;
    }

    vec_u32 bidders = {0}, objects_bidded = {0};
    vec_f64 bids = {0};
    for (;;) {
        forward_bid_and_assign(s, z, &bidders, &objects_bidded, &bids);
        s->nits += 1;
        if (z->num_unassigned == 0) {
            int is_optimal = start_from_optimal_eps ||
                             orc_ecs_satisfied(s, z->person_to_object.p, target_eps, toleration);
            if (is_optimal) {
                s->optimal_soln_found = 1;
                break;
            } else {
                if (z->eps < target_eps) break;
                z->eps *= 0.15; /* REDUCTION_FACTOR, symmetric.rs:189 */
                for (size_t i = 0; i < z->person_to_object.len; ++i) z->person_to_object.p[i] = s->imax;
                for (size_t j = 0; j < z->object_to_person.len; ++j) z->object_to_person.p[j] = s->imax;
                z->num_unassigned = s->num_rows;
                for (size_t i = 0; i < s->unassigned_people.len; ++i) s->unassigned_people.p[i] = (uint32_t)i;
                for (size_t i = 0; i < s->person_to_assignment_idx.len; ++i)
                    s->person_to_assignment_idx.p[i] = (uint32_t)i;
                s->nreductions += 1;
            }
        }
        if (s->nits >= s->max_iterations) break;
    }
    free(bidders.p); free(objects_bidded.p); free(bids.p);
    return ORC_OK;
}

/* ---- accessors for the ctypes wrapper ---- */
uint32_t orc_num_rows(const orc_solver *s) { return s->num_rows; }
uint32_t orc_num_cols(const orc_solver *s) { return s->num_cols; }
uint32_t orc_nits(const orc_solver *s) { return s->nits; }
uint32_t orc_nreductions(const orc_solver *s) { return s->nreductions; }
int orc_optimal_soln_found(const orc_solver *s) { return s->optimal_soln_found; }
uint64_t orc_bid_arcs(const orc_solver *s) { return s->bid_arcs; }
const double *orc_prices(const orc_solver *s) { return s->prices.p; }
size_t orc_prices_len(const orc_solver *s) { return s->prices.len; }
const double *orc_values(const orc_solver *s) { return s->values.p; }
const uint32_t *orc_column_indices(const orc_solver *s) { return s->column_indices.p; }
const uint32_t *orc_i_starts_stops(const orc_solver *s) { return s->i_starts_stops.p; }
size_t orc_i_starts_stops_len(const orc_solver *s) { return s->i_starts_stops.len; }
const uint32_t *orc_j_counts(const orc_solver *s) { return s->j_counts.p; }
size_t orc_j_counts_len(const orc_solver *s) { return s->j_counts.len; }
const uint32_t *orc_person_to_object(const orc_solution *z) { return z->person_to_object.p; }
size_t orc_person_to_object_len(const orc_solution *z) { return z->person_to_object.len; }
const uint32_t *orc_object_to_person(const orc_solution *z) { return z->object_to_person.p; }
size_t orc_object_to_person_len(const orc_solution *z) { return z->object_to_person.len; }
uint32_t orc_num_unassigned(const orc_solution *z) { return z->num_unassigned; }
double orc_solution_eps(const orc_solution *z) { return z->eps; }
