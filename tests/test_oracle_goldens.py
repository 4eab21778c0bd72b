"""CPU-only: pins the oracle (oracle/sla_oracle.c + oracle/fixture_rng.c) against every known-answer value the
reference's own tests hold for the hot path, and against scipy as an independent optimum."""
import numpy as np
import pytest

from helpers import U32_MAX, check_matching, dense_csr, fixtures, goldens, objective_of, random_sparse_instance, scipy_optimum

KINDS = ("khosla", "forward")


def solve(O, kind, n, m, row_ptr, cols, vals, **kw):
    s = O.OracleSolver(kind, n, m, len(cols))
    s.load_csr(n, m, row_ptr, cols, vals)
    s.solve(**kw)
    return s


def test_fixture_rng_matches_committed_arrays(oracle):
    fx = fixtures()
    for name, (n, m, k) in {"small": (5, 5, 2), "no_perfect": (9, 9, 3), "large": (90, 900, 32)}.items():
        rp, c, v = oracle.fixture_ksparse(n, m, k, 10.0)
        assert np.array_equal(rp, fx[name + "_row_ptr"])
        assert np.array_equal(c, fx[name + "_cols"])
        assert np.array_equal(v, fx[name + "_vals"])          # bit-exact f64
    assert np.array_equal(oracle.chacha8_u64_stream(1, 40), fx["chacha8_seed1_u64"])


@pytest.mark.parametrize("kind", KINDS)
def test_random_solve_small(oracle, kind):
    g = goldens()["random_solve_small"]                       # src/solver.rs:294-315
    csr = oracle.fixture_ksparse(g["rows"], g["cols"], g["k"], g["max_value"])
    for maximize, key in ((False, "minimize"), (True, "maximize")):
        s = solve(oracle, kind, g["rows"], g["cols"], *csr, maximize=maximize)
        assert s.get_objective() == g[key]                    # assert_eq! on f64: bit-exact
        assert s.num_unassigned == 0


@pytest.mark.parametrize("kind", KINDS)
def test_random_no_perfect_matching(oracle, kind):
    g = goldens()["random_no_perfect_matching"]               # src/solver.rs:317-337
    csr = oracle.fixture_ksparse(g["rows"], g["cols"], g["k"], g["max_value"])
    s = solve(oracle, kind, g["rows"], g["cols"], *csr)
    assert s.num_unassigned == 1
    assert s.get_objective() in g["accepted_objectives"]
    expected = g["accepted_objectives"][0 if kind == "khosla" else 1]
    assert s.get_objective() == expected


@pytest.mark.parametrize("kind", KINDS)
def test_random_large(oracle, kind):
    g = goldens()["random_large"]                             # src/solver.rs:419-437
    csr = oracle.fixture_ksparse(g["rows"], g["cols"], g["k"], g["max_value"])
    s = solve(oracle, kind, g["rows"], g["cols"], *csr)
    assert s.get_objective() == g["minimize"]
    assert s.num_unassigned == 0


@pytest.mark.parametrize("kind", KINDS)
def test_fixed_cases_reusing_one_solver(oracle, kind):
    g = goldens()                                             # src/solver.rs:339-418
    s = oracle.OracleSolver(kind, 10, 10, 100)                # one solver re-initialised for all cases (390-406)
    internals = g["probe_internals"]
    for idx, case in enumerate(g["fixed_cases"]["cases"]):
        s.load_dense(case["costs"])
        s.solve(maximize=False)
        assert s.num_unassigned == 0
        assert s.get_objective() == case["objective"]
        assert list(s.person_to_object) == case["person_to_object"]
        assert list(s.object_to_person) == case["object_to_person"]
        if kind == "khosla":
            assert s.nits == internals["khosla_nits"]["fixed"][idx]
        else:
            assert [s.nits, s.nreductions] == internals["forward_nits_nreductions"]["fixed"][idx]
            if idx == 0:
                assert s.eps == internals["forward_eps_8x8"]


@pytest.mark.parametrize("kind", KINDS)
def test_doctest_ragged(oracle, kind):
    g = goldens()["doctest"]                                  # src/ksparse.rs:22-72
    s = oracle.OracleSolver(kind, 10, 10, 100)
    s.init(2, 4)
    for i, row in enumerate(g["rows"]):
        s.extend_from_values(i, np.arange(len(row)), row)
    s.solve(maximize=False)
    assert s.num_unassigned == 0
    assert s.get_objective() == g["objective"]
    assert list(s.person_to_object) == g["person_to_object"]
    assert list(s.object_to_person) == g["object_to_person"]


def test_push_all_left(oracle):
    g = goldens()["push_all_left"]                            # src/symmetric.rs:516-523
    data, _ = oracle.push_all_left(g["data"], g["mapper"], g["num_ints"], g["size"], imax=0xFFFF)
    assert list(data) == g["expected"]


def test_cumulative_idx_diff_u16(oracle):
    g = goldens()["cumulative_idx_diff"]                      # src/symmetric.rs:525-534
    s = oracle.OracleSolver("forward", 7, 7, 7, imax=0xFFFF)
    s.init(7, 7)
    for r in g["rows"]:
        s.add_value(r, 0, 0.0)
    assert list(s.i_starts_stops) == g["i_starts_stops"]
    assert list(s.j_counts) == g["j_counts"]


def test_builder_errors(oracle):
    s = oracle.OracleSolver("khosla", 4, 4, 16)
    with pytest.raises(oracle.OracleError):
        s.init(5, 4)                                          # rows <= cols, solver.rs:192
    s.init(2, 4)
    with pytest.raises(oracle.OracleError):
        s.add_value(1, 0, 1.0)                                # row 0 still empty, solver.rs:55
    s.add_value(0, 0, 1.0)
    with pytest.raises(oracle.OracleError):
        s.add_value(2, 0, 1.0)                                # skipping a row, solver.rs:44
    s16 = oracle.OracleSolver("khosla", 4, 4, 16, imax=0xFFFF)
    with pytest.raises(oracle.OracleError):
        s16.init(0xFFFF, 0xFFFF)                              # num_rows < I::MAX, solver.rs:193
    empty = oracle.OracleSolver("forward", 4, 4, 16)
    empty.init(2, 2)
    with pytest.raises(oracle.OracleError):
        empty.solve()                                         # validate_input: no arcs, solver.rs:234


def test_toleration(oracle):
    assert oracle.get_toleration(1000.0) == 1.0 / 2 ** (53 - 9)    # solver.rs:144-146
    assert oracle.get_toleration(0.0) == 1.0 / 2 ** 53
    assert oracle.get_toleration(10.0) == 1.0 / 2 ** 50


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("seed", range(6))
def test_integer_instances_match_scipy_optimum(oracle, kind, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(4, 40))
    m = n if kind == "forward" or seed % 2 else n + int(rng.integers(1, 20))
    k = int(rng.integers(2, min(8, m) + 1))
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True)
    maximize = bool(seed % 2)
    s = solve(oracle, kind, n, m, rp, c, v, maximize=maximize, eps=1.0 / (m + 1))
    check_matching(n, m, rp, c, s.person_to_object, s.object_to_person, s.num_unassigned)
    assert s.num_unassigned == 0
    assert objective_of(rp, c, v, s.person_to_object) == scipy_optimum(n, m, rp, c, v, maximize)
    assert s.get_objective() == abs(objective_of(rp, c, v, s.person_to_object))


@pytest.mark.parametrize("algo", KINDS)
def test_jacobi_model_agrees_with_oracle(oracle, algo):
    """The CPU model of the device algorithm reaches the reference's objective / num_unassigned on every fixture."""
    for (n, m, k) in ((5, 5, 2), (9, 9, 3), (90, 900, 32)):
        rp, c, v = oracle.fixture_ksparse(n, m, k, 10.0)
        for maximize in (False, True):
            r = oracle.jacobi_model(algo, n, m, rp, c, v, maximize=maximize)
            s = solve(oracle, algo, n, m, rp, c, v, maximize=maximize)
            assert r["stats"]["num_unassigned"] == s.num_unassigned
            check_matching(n, m, rp, c, r["p2o"], r["o2p"], r["stats"]["num_unassigned"])
            assert abs(objective_of(rp, c, v, r["p2o"])) == s.get_objective()
    for case in goldens()["fixed_cases"]["cases"]:
        n, m, rp, c, v = dense_csr(case["costs"])
        r = oracle.jacobi_model(algo, n, m, rp, c, v)
        assert objective_of(rp, c, v, r["p2o"]) == case["objective"]
        assert r["stats"]["num_unassigned"] == 0


def test_model_khosla_eps_schedule_keeps_parity_with_the_oracle():
    """The device's Khosla rounds on square instances run under an eps-schedule (oracle/jacobi_model.c
    is the CPU model of exactly that).  Against the restated reference (sla_oracle.c): same num_unassigned, objective
    within n * eps (identical for integer weights with eps < 1/n); an instance without a perfect matching falls back to
    the plain rounds and returns what they return."""
    from oracle import oracle as O

    def instance(n, mean_degree, seed, planted, lo, hi, integer=False):
        rng = np.random.default_rng(seed)
        perm = rng.permutation(n)
        rows = []
        for i in range(n):
            cc = rng.choice(n, size=max(int(rng.binomial(n, mean_degree / n)), 1), replace=False)
            rows.append(np.unique(np.append(cc, perm[i]) if planted else cc))
        rp = np.zeros(n + 1, dtype=np.uint32)
        rp[1:] = np.cumsum([len(r) for r in rows])
        c = np.concatenate(rows).astype(np.uint32)
        v = rng.uniform(lo, hi, size=c.size)
        return rp, c, (np.floor(v) if integer else v)

    def objective(rp, c, v, p2o):
        total = 0.0
        for i, j in enumerate(p2o):
            if j != 0xFFFFFFFF:
                a, b = int(rp[i]), int(rp[i + 1])
                total += v[a + int(np.nonzero(c[a:b] == j)[0][0])]
        return total

    for n, seed, integer in ((200, 1, False), (200, 2, True), (600, 3, False)):
        rp, c, v = instance(n, 10, seed, True, 500.0, 1000.0, integer)
        eps = 1.0 / (n + 1) if integer else None
        o = O.OracleSolver("khosla", n, n, len(c))
        o.load_csr(n, n, rp, c, v)
        o.solve(eps=eps)
        scaled = O.jacobi_model("khosla", n, n, rp, c, v, eps=eps)
        plain = O.jacobi_model("khosla", n, n, rp, c, v, eps=eps, khosla_scaling=False)
        assert scaled["stats"]["num_unassigned"] == plain["stats"]["num_unassigned"] == o.num_unassigned == 0
        assert scaled["stats"]["nreductions"] >= 4 and scaled["stats"]["eps"] == o.eps
        got = objective(rp, c, v, scaled["p2o"])
        assert got == o.get_objective() if integer else abs(got - o.get_objective()) <= n * o.eps
    fallbacks = 0
    for n, seed in ((9, 5), (60, 6), (40, 8), (150, 7)):
        rp, c, v = instance(n, 3, seed, False, 0.0, 10.0)
        o = O.OracleSolver("khosla", n, n, len(c))
        o.load_csr(n, n, rp, c, v)
        o.solve()
        scaled = O.jacobi_model("khosla", n, n, rp, c, v)
        assert scaled["stats"]["num_unassigned"] == o.num_unassigned
        if o.num_unassigned:
            fallbacks += 1
            plain = O.jacobi_model("khosla", n, n, rp, c, v, khosla_scaling=False)
            assert np.array_equal(scaled["p2o"], plain["p2o"]) and np.array_equal(scaled["prices"], plain["prices"])
    assert fallbacks >= 2


def test_model_reports_restarts_only_without_a_perfect_matching(oracle):
    """`restarts` of the CPU model (the counter sla_stats.restarts is compared with on the GPU): a Khosla solve on a
    square instance abandons its eps-schedule exactly when somebody is dropped at the price threshold.  On the tests'
    instance families no feasible instance does, every infeasible one does (the full statistics:
    tests/khosla_restart_stats.py -> profiles/r02_khosla_restart_stats.jsonl), and a rectangular or unscheduled solve
    never reports one."""
    import khosla_restart_stats as K

    def planted(seed):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(100, 400))
        return (n,) + K.symmetric_instance(n, float(rng.uniform(2.0, 4.0)), 800 + seed, True, 0.0, 10.0)

    def unplanted(seed):
        rng = np.random.default_rng(seed)
        n = int(rng.integers(10, 120))
        return (n,) + K.symmetric_instance(n, float(rng.uniform(2.0, 4.0)), 900 + seed, False, 0.0, 10.0)

    a = K.family("planted", planted, 12)
    assert a["feasible"] == a["instances"] and a["restarts_among_all"] == 0
    assert a["rounds_feasible_with_schedule"] * 2 < a["rounds_feasible_plain"]
    b = K.family("unplanted", unplanted, 40)
    assert b["feasible"] < b["instances"]
    assert b["restarts_among_feasible"] == 0 and b["restarts_among_all"] == b["instances"] - b["feasible"]
    n, rp, c, v = unplanted(3)
    assert oracle.jacobi_model("khosla", n, n, rp, c, v, khosla_scaling=False)["stats"]["restarts"] == 0
    from helpers import random_sparse_instance
    rp, c, v = random_sparse_instance(np.random.default_rng(5), 50, 80, 6, integer=True, lo=0, hi=100)
    assert oracle.jacobi_model("khosla", 50, 80, rp, c, v)["stats"]["restarts"] == 0
