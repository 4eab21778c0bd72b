// test_api.cpp -- the reference's own test-suite (src/solver.rs:246-445, src/symmetric.rs:510-535, the doctest
// src/ksparse.rs:22-72) replayed through the C++ host mirror include/sla.hpp over the C ABI.
//   test_api host                 host-only checks (no GPU needed): CSR builder, errors, loud failure without a device
//   test_api gpu <fixtures.bin>   everything, on a B200; fixtures.bin holds the reference's three random CSR inputs
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <vector>

#include "sla.hpp"

static int failures = 0;
#define CHECK(cond)                                                                  \
    do {                                                                             \
        if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
    } while (0)

template <class F>
static bool throws(F f, int code = -1) {
    try { f(); } catch (const sla::Error& e) { return code < 0 || e.code == code; }
    return false;
}

struct Csr { uint32_t n, m; std::vector<uint32_t> rp, c; std::vector<double> v; };

static std::vector<Csr> read_fixtures(const char* path) {
    std::ifstream f(path, std::ios::binary);
    std::vector<Csr> out;
    uint32_t count = 0;
    f.read((char*)&count, 4);
    for (uint32_t i = 0; i < count; ++i) {
        Csr x; uint32_t nnz = 0;
        f.read((char*)&x.n, 4); f.read((char*)&x.m, 4); f.read((char*)&nnz, 4);
        x.rp.resize(x.n + 1); x.c.resize(nnz); x.v.resize(nnz);
        f.read((char*)x.rp.data(), 4 * (x.n + 1)); f.read((char*)x.c.data(), 4 * nnz); f.read((char*)x.v.data(), 8 * nnz);
        out.push_back(std::move(x));
    }
    return out;
}

template <class Solver>
static void populate(Solver& s, const Csr& x) {
    s.init(x.n, x.m);
    for (uint32_t i = 0; i < x.n; ++i)
        s.extend_from_values(i, x.c.data() + x.rp[i], x.rp[i + 1] - x.rp[i], x.v.data() + x.rp[i], x.rp[i + 1] - x.rp[i]);
}

template <class Solver>
static void populate_dense(Solver& s, const std::vector<std::vector<int>>& costs) {
    s.init(uint32_t(costs.size()), uint32_t(costs[0].size()));
    for (uint32_t i = 0; i < costs.size(); ++i) {
        std::vector<uint32_t> j(costs[i].size());
        std::vector<double> v(costs[i].size());
        for (size_t t = 0; t < j.size(); ++t) { j[t] = uint32_t(t); v[t] = double(costs[i][t]); }
        s.extend_from_values(i, j, v);
    }
}

static void host_tests() {
    {   // test_cumulative_idx_diff (symmetric.rs:525-534), u16 index type
        auto [solver, sol] = sla::ForwardAuctionSolver<uint16_t>::make(7, 7, 7);
        solver.init(7, 7);
        for (uint16_t r : {0, 0, 0, 1, 1, 1, 1}) solver.add_value(r, 0, 0.0);
        CHECK((solver.i_starts_stops() == std::vector<uint16_t>{0, 3, 7}));
        CHECK((solver.j_counts() == std::vector<uint16_t>{3, 4}));
        CHECK(sol.num_unassigned == 0xFFFF && std::isnan(sol.eps) && sol.person_to_object.empty());   // solution.rs:46-53
    }
    {   // builder / validation errors (solver.rs:44,55,75,192-193,234)
        auto [solver, sol] = sla::KhoslaSolver<uint32_t>::make(4, 4, 16);
        CHECK(throws([&] { solver.init(5, 4); }, SLA_ERR_INVALID));
        solver.init(2, 4);
        CHECK(throws([&] { solver.add_value(1, 0, 1.0); }, SLA_ERR_INVALID));
        solver.add_value(0, 0, 1.0);
        CHECK(throws([&] { solver.add_value(2, 0, 1.0); }, SLA_ERR_INVALID));
        CHECK(throws([&] { solver.extend_from_values(0, std::vector<uint32_t>{1, 2}, std::vector<double>{1.0}); }, SLA_ERR_INVALID));
        auto [s16, z16] = sla::KhoslaSolver<uint16_t>::make(4, 4, 16);
        CHECK(throws([&] { s16.init(0xFFFF, 0xFFFF); }, SLA_ERR_INVALID));
        auto [empty, ze] = sla::ForwardAuctionSolver<uint32_t>::make(4, 4, 16);
        empty.init(2, 2);
        CHECK(throws([&] { empty.solve(ze, false); }, SLA_ERR_INVALID));   // validate_input before any device work
        CHECK(solver.get_toleration(1000.0) == 1.0 / double(1ull << 44));
        auto twin = solver;                                                // Clone
        twin.add_value(0, 1, 2.0);
        CHECK(solver.num_of_arcs() == 1 && twin.num_of_arcs() == 2);
    }
}

template <class Solver>
static void gpu_suite(const char* name, const std::vector<Csr>& fx, bool forward) {
    std::printf("-- %s\n", name);
    const uint32_t MAXU = 0xFFFFFFFFu;
    {   // test_random_solve_small (solver.rs:294-315)
        auto [solver, sol] = Solver::make(5, 5, 10);
        const double golden[2] = {19.329346102942907, 26.682897194725648};
        for (int maximize = 0; maximize < 2; ++maximize) {
            populate(solver, fx[0]);
            solver.solve(sol, maximize != 0);
            CHECK(solver.get_objective(sol) == golden[maximize]);
            CHECK(sol.num_unassigned == 0);
        }
    }
    {   // test_random_no_perfect_matching (solver.rs:317-337)
        auto [solver, sol] = Solver::make(9, 9, 27);
        populate(solver, fx[1]);
        solver.solve(sol, false);
        CHECK(sol.num_unassigned == 1);
        const double obj = solver.get_objective(sol);
        CHECK(obj == 19.00601422087291 || obj == 27.812843918178544);
    }
    {   // test_random_large (solver.rs:419-437)
        auto [solver, sol] = Solver::make(90, 900, 90 * 32);
        populate(solver, fx[2]);
        solver.solve(sol, false);
        CHECK(solver.get_objective(sol) == 32.48411883859272);
        CHECK(sol.num_unassigned == 0);
    }
    {   // test_fixed_cases (solver.rs:339-418): one solver re-initialised for every case
        auto [solver, sol] = Solver::make(10, 10, 100);
        struct Case { std::vector<std::vector<int>> costs; double obj; std::vector<uint32_t> p2o, o2p; };
        std::vector<Case> cases = {
            {{{1000, 2, 11, 10, 8, 7, 6, 5}, {6, 1000, 1, 8, 8, 4, 6, 7}, {5, 12, 1000, 11, 8, 12, 3, 11}, {11, 9, 10, 1000, 1, 9, 8, 10},
              {11, 11, 9, 4, 1000, 2, 10, 9}, {12, 8, 5, 2, 11, 1000, 11, 9}, {10, 11, 12, 10, 9, 12, 1000, 3}, {10, 10, 10, 10, 6, 3, 1, 1000}},
             17.0, {1, 2, 0, 4, 5, 3, 7, 6}, {2, 0, 1, 5, 3, 4, 7, 6}},
            {{{10, 10, 13}, {4, 8, 8}, {8, 5, 8}}, 22.0, {1, 0, 2}, {1, 0, 2}},
            {{{10, 6, 14, 1}, {17, 18, 17, 15}, {14, 17, 15, 8}, {11, 13, 11, 4}}, 41.0, {1, 2, 0, 3}, {2, 0, 1, 3}},
            {{{10, 6, 14, 1}}, 1.0, {3}, {MAXU, MAXU, MAXU, 0}},
        };
        for (auto& cs : cases) {
            populate_dense(solver, cs.costs);
            solver.solve(sol, false);
            CHECK(sol.num_unassigned == 0);
            CHECK(solver.get_objective(sol) == cs.obj);
            if (forward) {   // same Jacobi order as the reference: the exact vectors of solver.rs:361-386
                CHECK(sol.person_to_object == cs.p2o);
                CHECK(sol.object_to_person == cs.o2p);
            }
        }
    }
    {   // doctest (ksparse.rs:22-72): ragged 2 x 4
        auto [solver, sol] = Solver::make(10, 10, 100);
        solver.init(2, 4);
        solver.extend_from_values(0, std::vector<uint32_t>{0, 1, 2, 3}, std::vector<double>{10, 6, 14, 1});
        solver.extend_from_values(1, std::vector<uint32_t>{0, 1, 2}, std::vector<double>{17, 18, 16});
        solver.solve(sol, false);
        CHECK(sol.num_unassigned == 0);
        CHECK(solver.get_objective(sol) == 17.0);
        CHECK((sol.person_to_object == std::vector<uint32_t>{3, 2}));
        CHECK((sol.object_to_person == std::vector<uint32_t>{MAXU, MAXU, 1, 0}));
    }
}

int main(int argc, char** argv) {
    const bool gpu = argc > 1 && std::strcmp(argv[1], "gpu") == 0;
    host_tests();
    if (!gpu) {
        // no device: the product path must fail loudly (SLA_ERR_NO_DEVICE), never fall back to a CPU solve
        auto [solver, sol] = sla::KhoslaSolver<uint32_t>::make(4, 4, 16);
        solver.init(2, 2);
        solver.extend_from_values(0, std::vector<uint32_t>{0, 1}, std::vector<double>{1.0, 2.0});
        solver.extend_from_values(1, std::vector<uint32_t>{0, 1}, std::vector<double>{3.0, 1.0});
        bool ok = false;
        try { solver.solve(sol, false); ok = true; } catch (const sla::Error& e) { CHECK(e.code == SLA_ERR_NO_DEVICE); }
        if (ok) std::printf("note: a GPU is present, solve succeeded (objective %.1f)\n", solver.get_objective(sol));
    } else {
        if (argc < 3) { std::printf("usage: test_api gpu fixtures.bin\n"); return 2; }
        const auto fx = read_fixtures(argv[2]);
        if (fx.size() != 3) { std::printf("bad fixture file\n"); return 2; }
        gpu_suite<sla::KhoslaSolver<uint32_t>>("KhoslaSolver<u32>", fx, false);
        gpu_suite<sla::ForwardAuctionSolver<uint32_t>>("ForwardAuctionSolver<u32>", fx, true);
        {   // u16 instantiation end to end + solve_with_params counters
            auto [solver, sol] = sla::ForwardAuctionSolver<uint16_t>::make(8, 8, 64);
            solver.init(3, 3);
            const int c[3][3] = {{10, 10, 13}, {4, 8, 8}, {8, 5, 8}};
            for (uint16_t i = 0; i < 3; ++i)
                solver.extend_from_values(i, std::vector<uint16_t>{0, 1, 2}, std::vector<double>{double(c[i][0]), double(c[i][1]), double(c[i][2])});
            solver.solve_with_params(sol, false, std::nullopt, std::nullopt, std::nullopt);
            CHECK(solver.get_objective(sol) == 22.0 && sol.num_unassigned == 0);
            CHECK(solver.nits == 9 && solver.nreductions == 2 && solver.optimal_soln_found);
            CHECK((sol.person_to_object == std::vector<uint16_t>{1, 0, 2}));
            CHECK(solver.values()[0] == -10.0);                     // in-place sign normalisation (solver.rs:214-216)
            CHECK(solver.ecs_satisfied(sol.person_to_object, 1.0 / 3.0, solver.get_toleration(13.0)));
        }
    }
    std::printf(failures ? "%d FAILURES\n" : "all C++ API checks passed\n", failures);
    return failures ? 1 : 0;
}
