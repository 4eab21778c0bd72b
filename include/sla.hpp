// sla.hpp -- header-only C++17 host mirror of the reference's public API over the C ABI (include/sla.h).
//
// The reference crate is Rust and there is no Rust toolchain in the build image, so this header is the compiled-code
// stand-in of the Rust shim: same names, argument meaning, post-conditions and error behaviour as
//   AuctionSolver trait           reference src/solver.rs:8-244
//   AuctionSolution<I>            reference src/solution.rs:22-53
//   KhoslaSolver<I>               reference src/ksparse.rs:73-260
//   ForwardAuctionSolver<I>       reference src/symmetric.rs:75-508
// The host keeps the CSR (std::vector, like the reference's Vec), mirrors it into HBM when it changed and calls the
// CUDA path.  `Result` = anyhow::Result<()>: failures throw sla::Error (nothing here computes an auction on the CPU).
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <limits>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "sla.h"

namespace sla {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("sla error " + std::to_string(c) + ": " + m), code(c) {}
};

inline void ensure(bool cond, const char* what) {   // anyhow::ensure!
    if (!cond) throw Error(SLA_ERR_INVALID, what);
}

/// reference src/solution.rs:22-53
template <class I>
struct AuctionSolution {
    static_assert(std::is_same<I, uint16_t>::value || std::is_same<I, uint32_t>::value, "UnsignedInt is u16 or u32");
    std::vector<I> person_to_object, object_to_person;
    I num_unassigned = std::numeric_limits<I>::max();
    double eps = std::numeric_limits<double>::quiet_NaN();
    AuctionSolution(size_t row_capacity = 0, size_t column_capacity = 0) {
        person_to_object.reserve(row_capacity);
        object_to_person.reserve(column_capacity);
    }
};

/// Shared trait with default methods (reference src/solver.rs:8-244); CRTP is not needed because the only virtual
/// piece is solve().
template <class I>
class AuctionSolver {
public:
    static constexpr I IMAX = std::numeric_limits<I>::max();

    AuctionSolver(size_t row_capacity, size_t column_capacity, size_t arcs_capacity, int device = 0)
        : device_(device), caps_{row_capacity, column_capacity, arcs_capacity} {
        i_starts_stops_.reserve(row_capacity + 1);
        j_counts_.reserve(row_capacity);
        prices_.reserve(column_capacity);
        column_indices_.reserve(arcs_capacity);
        values_.reserve(arcs_capacity);
    }
    virtual ~AuctionSolver() { if (ctx_) sla_ctx_destroy(ctx_); }
    // #[derive(Clone)]: deep copy of the host state; the device context is re-created lazily
    AuctionSolver(const AuctionSolver& o)
        : nits(o.nits), num_rows_(o.num_rows_), num_cols_(o.num_cols_), prices_(o.prices_),
          i_starts_stops_(o.i_starts_stops_), j_counts_(o.j_counts_), column_indices_(o.column_indices_),
          values_(o.values_), device_(o.device_), caps_(o.caps_) {}
    AuctionSolver& operator=(const AuctionSolver&) = delete;

    // accessors (solver.rs:22-38); the *_mut twins mark the device mirror stale
    I num_rows() const { return num_rows_; }
    I num_cols() const { return num_cols_; }
    const std::vector<double>& prices() const { return prices_; }
    const std::vector<I>& i_starts_stops() const { return i_starts_stops_; }
    const std::vector<I>& j_counts() const { return j_counts_; }
    const std::vector<I>& column_indices() const { return column_indices_; }
    const std::vector<double>& values() const { return values_; }
    std::vector<double>& prices_mut() { return prices_; }
    std::vector<I>& i_starts_stops_mut() { dirty_ = true; return i_starts_stops_; }
    std::vector<I>& j_counts_mut() { dirty_ = true; return j_counts_; }
    std::vector<I>& column_indices_mut() { dirty_ = true; return column_indices_; }
    std::vector<double>& values_mut() { dirty_ = true; return values_; }

    /// solver.rs:191-205
    void init(I num_rows, I num_cols) {
        ensure(num_rows <= num_cols, "num_rows <= num_cols");
        ensure(num_rows < IMAX, "num_rows < I::MAX");
        num_rows_ = num_rows;
        num_cols_ = num_cols;
        i_starts_stops_.assign(2, I(0));
        j_counts_.assign(1, I(0));
        column_indices_.clear();
        values_.clear();
        dirty_ = true;
    }

    /// solver.rs:41-66
    void add_value(I row, I column, double value) {
        const size_t current_row = j_counts_.size() - 1;
        ensure(size_t(row) == current_row || size_t(row) == current_row + 1, "rows must arrive in non-decreasing order");
        const uint64_t cumulative = uint64_t(i_starts_stops_[current_row + 1]) + 1;
        ensure(cumulative <= IMAX, "i_starts_stops vector is longer then max value of type");
        if (size_t(row) > current_row) {
            ensure(j_counts_[current_row] > 0, "previous row is empty");
            i_starts_stops_.push_back(I(cumulative));
            j_counts_.push_back(I(1));
        } else {
            i_starts_stops_[current_row + 1] = I(cumulative);
            j_counts_[current_row] += 1;
        }
        column_indices_.push_back(column);
        values_.push_back(value);
        dirty_ = true;
    }

    /// solver.rs:69-101
    void extend_from_values(I row, const I* columns, size_t ncolumns, const double* values, size_t nvalues) {
        ensure(ncolumns == nvalues, "columns.len() == values.len()");
        const size_t current_row = j_counts_.size() - 1;
        ensure(size_t(row) == current_row || size_t(row) == current_row + 1, "rows must arrive in non-decreasing order");
        ensure(ncolumns <= size_t(IMAX), "columns slice is longer then max value of type");
        const uint64_t cumulative = uint64_t(i_starts_stops_[current_row + 1]) + ncolumns;
        ensure(cumulative <= IMAX, "i_starts_stops vector is longer then max value of type");
        if (size_t(row) > current_row) {
            ensure(j_counts_[current_row] > 0, "previous row is empty");
            i_starts_stops_.push_back(I(cumulative));
            j_counts_.push_back(I(ncolumns));
        } else {
            i_starts_stops_[current_row + 1] = I(cumulative);
            j_counts_[current_row] = I(j_counts_[current_row] + ncolumns);
        }
        column_indices_.insert(column_indices_.end(), columns, columns + ncolumns);
        values_.insert(values_.end(), values, values + nvalues);
        dirty_ = true;
    }
    void extend_from_values(I row, const std::vector<I>& columns, const std::vector<double>& values) {
        extend_from_values(row, columns.data(), columns.size(), values.data(), values.size());
    }

    size_t num_of_arcs() const { return column_indices_.size(); }   // solver.rs:104-106

    /// solver.rs:232-243
    void validate_input() const {
        const size_t arcs = num_of_arcs();
        ensure(arcs > 0, "arcs_count > 0");
        ensure(num_rows_ > 0 && num_cols_ > 0, "num_rows > 0 && num_cols > 0");
        ensure(arcs < size_t(IMAX), "arcs_count < I::MAX");
        ensure(arcs == column_indices_.size() && column_indices_.size() == values_.size(), "column_indices.len() == values.len()");
    }

    /// solver.rs:110-142 (sequential left-to-right sum: bit-identical to the reference for any weights)
    double get_objective(const AuctionSolution<I>& solution) const {
        const bool positive_values = (values_.empty() ? 0.0 : values_[0]) >= 0.0;
        double obj = 0.0;
        for (size_t i = 0; i < size_t(num_rows_); ++i) {
            const I j = solution.person_to_object[i];
            if (j == IMAX) continue;
            const size_t start = i_starts_stops_[i], n = j_counts_[i];
            for (size_t g = start; g < start + n; ++g)
                if (column_indices_[g] == j) obj += positive_values ? values_[g] : -values_[g];
        }
        return obj;
    }

    /// solver.rs:144-146
    double get_toleration(double max_abs_cost) const {
        const double l = std::log2(max_abs_cost + 1e-7);
        const uint32_t li = !(l > 0.0) ? 0u : (l >= 4294967295.0 ? 4294967295u : uint32_t(l));
        const uint32_t e = 53u - li;
        const uint64_t pw = e < 64 ? (uint64_t(1) << e) : 0;
        return 1.0 / double(pw);
    }

    /// solver.rs:154-189 on the host copies
    bool ecs_satisfied(const std::vector<I>& person_to_object, double eps, double toleration) const {
        for (size_t i = 0; i < size_t(num_rows_); ++i) {
            const size_t start = i_starts_stops_[i], n = j_counts_[i];
            const I j = person_to_object[i];
            double chosen = -std::numeric_limits<double>::infinity();
            for (size_t g = start; g < start + n; ++g)
                if (column_indices_[g] == j) chosen = values_[g];
            const double lhs = chosen - prices_.at(size_t(j)) + toleration;
            for (size_t g = start; g < start + n; ++g)
                if (lhs < values_[g] - prices_[size_t(column_indices_[g])] - eps) return false;
        }
        return true;
    }

    virtual void solve(AuctionSolution<I>& solution, bool maximize, std::optional<double> eps = std::nullopt) = 0;

    uint32_t nits = 0;
    sla_stats last_stats{};

protected:
    void check(int rc) const {
        if (rc != SLA_OK) throw Error(rc, sla_last_error(ctx_) ? sla_last_error(ctx_) : "");
    }
    sla_ctx* context() {
        if (!ctx_) {
            int rc = sla_ctx_create(device_, caps_[0], caps_[1], caps_[2], &ctx_);
            if (rc != SLA_OK) { ctx_ = nullptr; throw Error(rc, sla_last_error(nullptr)); }
            dirty_ = true;
        }
        return ctx_;
    }
    sla_ctx* sync_device() {
        sla_ctx* ctx = context();
        if (dirty_) {
            const size_t n = size_t(num_rows_), nnz = num_of_arcs();
            ensure(i_starts_stops_.size() >= n + 1, "fewer rows populated than num_rows");
            if constexpr (std::is_same<I, uint32_t>::value) {
                check(sla_upload_csr(ctx, num_rows_, num_cols_, i_starts_stops_.data(), column_indices_.data(), values_.data(), nnz));
            } else {   // u16 instantiation: widen the indices
                std::vector<uint32_t> rp(i_starts_stops_.begin(), i_starts_stops_.begin() + n + 1);
                std::vector<uint32_t> ci(column_indices_.begin(), column_indices_.end());
                check(sla_upload_csr(ctx, num_rows_, num_cols_, rp.data(), ci.data(), values_.data(), nnz));
            }
            dirty_ = false;
        }
        return ctx;
    }
    // copies the results into the caller's solution (u32::MAX truncates to u16::MAX) and applies the observable
    // in-place sign normalisation of `values` (solver.rs:214-216)
    void finish(AuctionSolution<I>& solution, const sla_stats& st, const std::vector<uint32_t>& p2o, const std::vector<uint32_t>& o2p) {
        if (st.values_negated) sla_host_negate_f64(values_.data(), values_.size(), 8);
        solution.person_to_object.assign(p2o.begin(), p2o.end());
        solution.object_to_person.assign(o2p.begin(), o2p.end());
        if constexpr (!std::is_same<I, uint32_t>::value) {
            for (size_t i = 0; i < p2o.size(); ++i) solution.person_to_object[i] = I(p2o[i]);
            for (size_t j = 0; j < o2p.size(); ++j) solution.object_to_person[j] = I(o2p[j]);
        }
        solution.num_unassigned = I(st.num_unassigned);
        solution.eps = st.eps;
        nits = st.nits;
        last_stats = st;
    }

    I num_rows_ = 0, num_cols_ = 0;
    std::vector<double> prices_;
    std::vector<I> i_starts_stops_, j_counts_, column_indices_;
    std::vector<double> values_;
    int device_ = 0;
    std::array<size_t, 3> caps_;
    sla_ctx* ctx_ = nullptr;
    bool dirty_ = true;
};

/// reference src/ksparse.rs:73-260
template <class I = uint32_t>
class KhoslaSolver : public AuctionSolver<I> {
public:
    using Base = AuctionSolver<I>;
    using Base::Base;
    /// solver.rs:9-13
    static std::pair<KhoslaSolver, AuctionSolution<I>> make(size_t row_capacity, size_t column_capacity, size_t arcs_capacity) {
        return {KhoslaSolver(row_capacity, column_capacity, arcs_capacity), AuctionSolution<I>(row_capacity, column_capacity)};
    }
    /// ksparse.rs:153-251
    void solve(AuctionSolution<I>& solution, bool maximize, std::optional<double> eps = std::nullopt) override {
        this->validate_input();
        sla_ctx* ctx = this->sync_device();
        std::vector<uint32_t> p2o(size_t(this->num_rows_)), o2p(size_t(this->num_cols_));
        this->prices_.assign(size_t(this->num_cols_), 0.0);
        sla_stats st{};
        this->check(sla_khosla_solve(ctx, maximize ? 1 : 0, eps ? *eps : std::numeric_limits<double>::quiet_NaN(), p2o.data(),
                                     o2p.data(), this->prices_.data(), &st));
        this->finish(solution, st, p2o, o2p);
    }
};

/// reference src/symmetric.rs:75-508
template <class I = uint32_t>
class ForwardAuctionSolver : public AuctionSolver<I> {
public:
    using Base = AuctionSolver<I>;
    using Base::Base;
    static constexpr double REDUCTION_FACTOR = 0.15;    // symmetric.rs:189
    static constexpr uint32_t MAX_ITERATIONS = 100000;  // symmetric.rs:190
    uint32_t max_iterations = MAX_ITERATIONS;
    uint32_t nreductions = 0;
    bool optimal_soln_found = false;

    static std::pair<ForwardAuctionSolver, AuctionSolution<I>> make(size_t row_capacity, size_t column_capacity, size_t arcs_capacity) {
        return {ForwardAuctionSolver(row_capacity, column_capacity, arcs_capacity), AuctionSolution<I>(row_capacity, column_capacity)};
    }
    /// symmetric.rs:177-185
    void solve(AuctionSolution<I>& solution, bool maximize, std::optional<double> eps = std::nullopt) override {
        solve_with_params(solution, maximize, eps, std::nullopt, std::nullopt);
    }
    /// symmetric.rs:217-332
    void solve_with_params(AuctionSolution<I>& solution, bool maximize, std::optional<double> eps, std::optional<double> start_eps,
                           std::optional<uint32_t> max_iter) {
        this->validate_input();
        sla_ctx* ctx = this->sync_device();
        std::vector<uint32_t> p2o(size_t(this->num_rows_)), o2p(size_t(this->num_cols_));
        this->prices_.assign(size_t(this->num_cols_), 0.0);
        // Some(0) behaves like Some(1) in the reference (the check runs after the first round, symmetric.rs:326)
        max_iterations = max_iter ? (*max_iter ? *max_iter : 1u) : MAX_ITERATIONS;
        const double nan = std::numeric_limits<double>::quiet_NaN();
        sla_stats st{};
        this->check(sla_forward_solve(ctx, maximize ? 1 : 0, eps ? *eps : nan, start_eps ? *start_eps : nan, max_iterations,
                                      p2o.data(), o2p.data(), this->prices_.data(), &st));
        this->finish(solution, st, p2o, o2p);
        nreductions = st.nreductions;
        optimal_soln_found = st.optimal_soln_found != 0;
    }
};

}  // namespace sla
