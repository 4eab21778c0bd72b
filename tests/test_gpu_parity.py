"""GPU parity tests (run with -m gpu on a B200).  Everything goes through the C ABI (libsla_b200.so) via the
host-side mirror of the reference API; the CPU oracle and the CPU model of the device algorithm are the checkers.

Bars (BASELINE.json north_star): valid matching with the reference's num_unassigned; bit-exact objective for
integer weights with eps < 1/n; within n * eps_final otherwise; and -- stronger, our own bar -- bit-exact equality
of prices / assignment vectors / counters with the sequential CPU model of the Jacobi rounds.
"""
import numpy as np
import pytest

from helpers import (U32_MAX, check_matching, dense_csr, fixtures, goldens, objective_of, random_sparse_instance,
                     scipy_optimum)

pytestmark = pytest.mark.gpu

SOLVERS = (("khosla", "KhoslaSolver"), ("forward", "ForwardAuctionSolver"))


def gpu_solve(sla, cls_name, n, m, rp, c, v, maximize=False, eps=None, options=None, index_dtype=np.uint32, **kw):
    solver, z = getattr(sla, cls_name).new(n, m, len(c), index_dtype=index_dtype)
    solver.load_csr(n, m, rp, c, v)
    for k_, v_ in (options or {}).items():
        solver.set_option(k_, v_)
    if kw:
        solver.solve_with_params(z, maximize, eps, kw.get("start_eps"), kw.get("max_iterations"))
    else:
        solver.solve(z, maximize, eps)
    return solver, z


def assert_equals_model(O, kind, solver, z, n, m, rp, c, v, maximize=False, eps=None, **kw):
    r = O.jacobi_model(kind, n, m, rp, c, v, maximize=maximize, eps=eps, **kw)
    st = solver.last_stats
    assert np.array_equal(z.person_to_object.astype(np.uint32), r["p2o"]), "person_to_object differs from the model"
    assert np.array_equal(z.object_to_person.astype(np.uint32), r["o2p"]), "object_to_person differs from the model"
    assert np.array_equal(solver.prices(), r["prices"]), "prices differ from the model (bit-exact f64)"
    for key in ("num_unassigned", "nits", "nreductions", "optimal_soln_found", "rounds", "bids", "bid_arcs", "dropped",
                "values_negated", "restarts"):
        assert st[key] == r["stats"][key], f"{key}: gpu {st[key]} != model {r['stats'][key]}"
    assert st["eps"] == r["stats"]["eps"]
    return r


def assert_reference_parity(O, kind, solver, z, n, m, rp, c, v, maximize=False, eps=None, integer=False):
    """The north star's bar against the REFERENCE (the pinned oracle, not the model of the device rounds): same
    num_unassigned; with everybody assigned the objective is bit-exact for integer weights when the final eps is
    below 1/n (both solutions are optimal) and within n * eps_final otherwise (both are within n * eps of the optimum:
    ksparse.rs / symmetric.rs end in eps-complementary slackness).  Returns the oracle."""
    if kind == "forward" and int(np.diff(np.asarray(rp, dtype=np.int64)).min()) < 2:
        return None      # single-arc rows make Forward bid +inf and then NaN (symmetric.rs:378, 394): no comparable end state
    o = O.OracleSolver(kind, n, m, len(c))
    o.load_csr(n, m, rp, c, v)
    o.solve(maximize=maximize, eps=eps)
    assert z.num_unassigned == o.num_unassigned, (kind, n, m, z.num_unassigned, o.num_unassigned)
    if z.num_unassigned == 0:
        got, ref = solver.get_objective(z), o.get_objective()
        e = max(z.eps, o.eps)
        if kind == "forward" and not (solver.optimal_soln_found and o.optimal_soln_found):
            return o     # cut short by max_iterations: a complete assignment, but no eps-CS certificate on one side
        if integer and e < 1.0 / n:
            assert got == ref, (kind, n, m, got, ref)
        else:
            assert abs(got - ref) <= n * e + 1e-9 * max(1.0, abs(ref)), (kind, n, m, got, ref, e)
    return o


# ---- the reference's own test-suite, instantiated for both solvers (src/solver.rs:246-445) ------------------------
@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_random_solve_small(sla, oracle, kind, cls_name):
    g = goldens()["random_solve_small"]
    fx = fixtures()
    rp, c, v = fx["small_row_ptr"], fx["small_cols"], fx["small_vals"]
    solver, z = getattr(sla, cls_name).new(5, 5, 10)
    for maximize, key in ((False, "minimize"), (True, "maximize")):
        solver.load_csr(5, 5, rp, c, v)                       # the reference re-populates before each solve
        solver.solve(z, maximize, None)
        assert solver.get_objective(z) == g[key]              # assert_eq! on f64
        assert z.num_unassigned == 0
        assert_equals_model(oracle, kind, solver, z, 5, 5, rp, c, v, maximize=maximize)


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_random_no_perfect_matching(sla, oracle, kind, cls_name):
    g = goldens()["random_no_perfect_matching"]
    fx = fixtures()
    rp, c, v = fx["no_perfect_row_ptr"], fx["no_perfect_cols"], fx["no_perfect_vals"]
    solver, z = gpu_solve(sla, cls_name, 9, 9, rp, c, v)
    assert z.num_unassigned == 1
    assert solver.get_objective(z) in g["accepted_objectives"]
    check_matching(9, 9, rp, c, z.person_to_object, z.object_to_person, z.num_unassigned)
    assert_equals_model(oracle, kind, solver, z, 9, 9, rp, c, v)
    if kind == "forward":
        assert solver.nits == 100000 and not solver.optimal_soln_found      # runs into MAX_ITERATIONS like the reference


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_random_large(sla, oracle, kind, cls_name):
    g = goldens()["random_large"]
    fx = fixtures()
    rp, c, v = fx["large_row_ptr"], fx["large_cols"], fx["large_vals"]
    solver, z = gpu_solve(sla, cls_name, 90, 900, rp, c, v)
    assert solver.get_objective(z) == g["minimize"]
    assert z.num_unassigned == 0
    assert_equals_model(oracle, kind, solver, z, 90, 900, rp, c, v)


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_fixed_cases_reusing_one_solver(sla, oracle, kind, cls_name):
    solver, z = getattr(sla, cls_name).new(10, 10, 100)
    for case in goldens()["fixed_cases"]["cases"]:
        n, m, rp, c, v = dense_csr(case["costs"])
        solver.init(n, m)
        for i in range(n):
            solver.extend_from_values(i, np.arange(m, dtype=np.uint32), np.asarray(case["costs"][i], dtype=np.float64))
        solver.solve(z, False, None)
        assert z.num_unassigned == 0
        assert solver.get_objective(z) == case["objective"]
        check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, 0)
        assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v)
        if kind == "forward":      # same Jacobi order as the reference and no ties: the exact vectors of solver.rs:361-386
            assert list(z.person_to_object) == case["person_to_object"]
            assert list(z.object_to_person) == case["object_to_person"]


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_doctest_ragged(sla, oracle, kind, cls_name):
    g = goldens()["doctest"]
    solver, z = getattr(sla, cls_name).new(10, 10, 100)
    solver.init(2, 4)
    for i, row in enumerate(g["rows"]):
        solver.extend_from_values(i, np.arange(len(row), dtype=np.uint32), np.asarray(row, dtype=np.float64))
    solver.solve(z, False, None)
    assert z.num_unassigned == 0
    assert solver.get_objective(z) == g["objective"]
    assert list(z.person_to_object) == g["person_to_object"]
    assert list(z.object_to_person) == g["object_to_person"]


# ---- seeded instances: GPU == model bit for bit, and objective parity with the reference oracle ------------------
def ragged_instance(rng, n, m, kmin, kmax, integer):
    """Rows of different lengths (unaligned row starts exercise the masked 128-bit chunk loads)."""
    counts = rng.integers(kmin, kmax + 1, size=n)
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum(counts)
    cols = np.empty(rp[-1], dtype=np.uint32)
    perm = rng.permutation(m)[:n]
    for i in range(n):
        k = int(counts[i])
        others = rng.choice(m - 1, size=k - 1, replace=False)
        others = others + (others >= perm[i])
        cols[rp[i]:rp[i + 1]] = np.sort(np.concatenate([[perm[i]], others]))
    vals = rng.integers(0, 1000, size=rp[-1]).astype(np.float64) if integer else rng.uniform(0, 10, size=rp[-1])
    return rp, cols, vals


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
@pytest.mark.parametrize("seed", range(8))
def test_seeded_instances_match_model_and_oracle(sla, oracle, kind, cls_name, seed):
    rng = np.random.default_rng(1000 + seed)
    integer = seed % 2 == 0
    n = int(rng.integers(20, 400))
    m = n if (kind == "forward" and seed % 4 < 2) else n + int(rng.integers(0, 200))
    kmax = int(rng.integers(3, 40))
    rp, c, v = ragged_instance(rng, n, m, 1 if seed == 3 else 2, min(kmax, m), integer)
    maximize = seed % 3 == 0
    eps = 1.0 / (m + 1) if integer else None
    solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v, maximize=maximize, eps=eps)
    assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v, maximize=maximize, eps=eps)
    check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, z.num_unassigned)
    o = oracle.OracleSolver(kind, n, m, len(c))
    o.load_csr(n, m, rp, c, v)
    o.solve(maximize=maximize, eps=eps)
    if seed == 3 and kind == "forward":
        return   # single-arc rows make Forward bid +inf (symmetric.rs:378): both sides only need to stay valid
    assert z.num_unassigned == o.num_unassigned
    if z.num_unassigned == 0:
        got, ref = solver.get_objective(z), o.get_objective()
        if integer:
            assert got == ref                                  # bit-exact: integer weights, eps < 1/n
            assert objective_of(rp, c, v, z.person_to_object) == scipy_optimum(n, m, rp, c, v, maximize)
        else:
            assert abs(got - ref) <= n * max(z.eps, o.eps) + 1e-9      # tolerance n * eps_final (north_star)
    assert np.array_equal(solver.values(), o.values)           # in-place sign normalisation is observable


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_engine_options_do_not_change_results(sla, oracle, kind, cls_name):
    """Wide kernels only / tail engine only / host loop / CUDA graph / gather skipped or not: identical bits."""
    rng = np.random.default_rng(7)
    n, m, k = 3000, (3000 if kind == "forward" else 5000), 16      # k % 8 == 0: the regular-CSR bid kernel applies
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=200)
    ref = oracle.jacobi_model(kind, n, m, rp, c, v, eps=1.0 / (m + 1))
    combos = [
        dict(graph=1, tail_max=1024, zero_price_skip=1),
        dict(graph=0, tail_max=1024, zero_price_skip=1),
        dict(graph=1, tail_max=0, zero_price_skip=0),
        dict(graph=0, tail_max=0, zero_price_skip=1, profile=1),
        dict(graph=1, tail_max=512, zero_price_skip=0, super_rounds=2),
        dict(graph=1, tail_max=64, zero_price_skip=1, super_rounds=16),
        dict(graph=1, tail_max=1024, zero_price_skip=1, regular=0),
        dict(graph=1, tail_max=1024, zero_price_skip=1, smem_prices=0),
        dict(graph=0, tail_max=16, zero_price_skip=0, regular=0),
        dict(graph=1, tail_max=1024, zero_price_skip=1, smem_owners=0),
        dict(graph=1, tail_max=300, zero_price_skip=1, smem_prices=0, smem_owners=0, regular=0),
        dict(graph=1, tail_max=1024, zero_price_skip=1, stream_scan=1),
        dict(graph=0, tail_max=0, zero_price_skip=1, stream_scan=1, profile=1),
    ]
    for opt in combos:
        solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v, eps=1.0 / (m + 1), options=opt)
        assert np.array_equal(z.person_to_object, ref["p2o"]), opt
        assert np.array_equal(z.object_to_person, ref["o2p"]), opt
        assert np.array_equal(solver.prices(), ref["prices"]), opt
        for key in ("num_unassigned", "nits", "nreductions", "rounds", "bids", "bid_arcs"):
            assert solver.last_stats[key] == ref["stats"][key], (opt, key)
        if opt.get("profile"):
            prof = solver.round_profile()
            assert prof and prof[0]["bidders"] == n and prof[0]["arcs"] == n * k
            assert sum(p["arcs"] for p in prof) == ref["stats"]["bid_arcs"]
        if opt["tail_max"] == 0:
            assert solver.last_stats["tail_rounds"] == 0


@pytest.mark.parametrize("k", [8, 16, 24, 40, 64, 136, 256])
@pytest.mark.parametrize("stream", [0, 1])
def test_first_round_scans_equal_model(sla, oracle, k, stream):
    """First-round scan of a regular CSR, both implementations -- the two-rows-in-flight LDG.256 kernel (default) and the
    TMA pipeline (bid_stream_kernel, option stream_scan): every lanes-per-row class, row counts that leave a ragged last
    pass / tile, both sign conventions; bit-identical to the sequential model."""
    rng = np.random.default_rng(100 + k)
    n = 2500 + k + 3                                   # not a multiple of any tile height
    m = 3 * n
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=500)
    for maximize in (False, True):
        solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, v, maximize=maximize, options=dict(stream_scan=stream, tail_max=256))
        assert solver.last_stats["wide_rounds"] >= 1
        assert_equals_model(oracle, "khosla", solver, z, n, m, rp, c, v, maximize=maximize)


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_small_instance_path_equals_device_statistics_path(sla, oracle, kind, cls_name):
    """Small instances take their statistics on the host while staging the upload and download their results behind
    the graph launch (option small_path, default on); the device-statistics path (small_path = 0) must see the same
    value range, sign, regularity and errors: ragged and regular rows, zeros of both signs, mixed signs, both
    objectives, repeated solves with the in-place negation observable on the host copy."""
    rng = np.random.default_rng(11)
    cases = []
    for n, m, k, regular in ((7, 7, 3, False), (60, 90, 8, True), (700, 2100, 16, True), (900, 900, 5, False), (1500, 4000, 32, True)):
        if kind == "forward":
            m = n if n in (7, 900) else m
        rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=50)
        if not regular:
            v = v - 20.0                      # mixed signs
            v[::7] = 0.0
            v[3::11] = -0.0
        cases.append((n, m, rp, c, v))
    for n, m, rp, c, v in cases:
        for maximize in (False, True):
            out = []
            for small in (1, 0):
                solver, z = getattr(sla, cls_name).new(n, m, len(c))
                solver.set_option("small_path", small)
                solver.load_csr(n, m, rp, c, v.copy())
                solver.solve(z, maximize, None)
                first = (z.person_to_object.copy(), z.object_to_person.copy(), solver.prices().copy(), solver.values().copy(),
                         dict(solver.last_stats))
                solver.solve(z, not maximize, None)          # second solve on the (possibly negated) host copy
                out.append((first, z.person_to_object.copy(), solver.prices().copy(), solver.values().copy()))
            a, b = out
            for x, y in zip(a[0][:4], b[0][:4]):
                assert np.array_equal(x, y)
            for key in ("num_unassigned", "nits", "nreductions", "rounds", "bids", "bid_arcs", "dropped", "values_negated", "eps"):
                assert a[0][4][key] == b[0][4][key], key
            for x, y in zip(a[1:], b[1:]):
                assert np.array_equal(x, y)
            solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v.copy(), maximize=maximize)
            assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v.copy(), maximize=maximize)


@pytest.mark.parametrize("flavour", ["u16", "f32", "u16_then_fraction", "u16_then_fraction_early", "u16_then_fraction_middle",
                                     "f64", "negative_ints", "minus_zero"])
def test_narrow_upload_is_lossless(sla, oracle, flavour):
    """Large uploads whose values survive a round trip through u16 / f32 cross PCIe narrow and are widened in HBM
    (sla_last_upload reports the width); anything else goes up as f64 -- also when the first non-representable value
    sits late in the array.  Either way the device sees the original bits: results equal those with narrow_upload = 0,
    and the caller's values end up negated in place exactly once."""
    rng = np.random.default_rng(5)
    n, m, k = 70_000, 200_000, 16                                   # 1.12 M arcs: above the narrowing threshold
    rp, c, v = sla.generators.kregular_host(n, m, k, seed=9)
    v = v.astype(np.float64)
    expect = {"u16": 2, "f32": 4, "u16_then_fraction": 8, "u16_then_fraction_early": 8, "u16_then_fraction_middle": 8,
              "f64": 8, "negative_ints": 4, "minus_zero": 4}[flavour]
    if flavour == "f32":
        v = v + 0.5
    elif flavour == "u16_then_fraction":
        v[-12345] += 0.1                                            # passes the probe, fails u16 (and f32) late
    elif flavour == "u16_then_fraction_early":
        v[70_001] += 0.1                                            # second 64 Ki-value piece: the copy loop has sent nothing yet
    elif flavour == "u16_then_fraction_middle":
        v[v.size // 2 + 3] += 0.1
    elif flavour == "f64":
        v = v + rng.uniform(0.0, 1.0, size=v.size)
    elif flavour == "negative_ints":
        v = -v                                                      # negative: not u16, exact in f32
    elif flavour == "minus_zero":
        v[5::97] = -0.0                                             # the sign of zero must survive: f32 keeps it
    out = []
    for narrow in (1, 0):
        solver, z = sla.KhoslaSolver.new(n, m, n * k)
        solver.set_option("narrow_upload", narrow)
        solver.load_csr(n, m, rp, c, v.copy())
        maximize = flavour == "negative_ints"                       # both cases need the sign flip (host negation)
        solver.solve(z, maximize, None)
        nbytes, width = solver.last_upload()
        assert width == (expect if narrow else 8), (flavour, narrow, width)
        assert nbytes == 4 * (n + 1) + 4 * n * k + width * n * k
        assert solver.last_stats["values_negated"] == 1
        assert np.array_equal(solver.values(), -v)                  # negated in place, exactly once
        out.append((z.person_to_object.copy(), z.object_to_person.copy(), solver.prices().copy(),
                    {q: solver.last_stats[q] for q in ("rounds", "bids", "bid_arcs", "num_unassigned", "eps")}))
    for x, y in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(x, y)
    assert out[0][3] == out[1][3]


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
@pytest.mark.parametrize("maximize", [False, True])
def test_narrow_scan_reads_the_u16_mirror_and_changes_nothing(sla, oracle, kind, cls_name, maximize):
    """After a u16 upload the grid-wide uniform-degree scans read the u16 copy of the values that stayed in HBM
    (sla_scan_value_bytes == 2, option narrow_scan): 6 instead of 12 bytes per arc.  The doubles are the same after the
    exact conversion, so prices / assignment / counters equal the f64 scan (narrow_scan = 0) and the CPU model bit for
    bit -- with and without the on-the-fly sign flip, with the extreme values 0 and 65,535 present, across a second
    solve on the resident CSR, and a later upload of other values drops the mirror."""
    n, m, k = (70_000, 200_000, 16) if kind == "khosla" else (66_000, 66_000, 16)
    rp, c, v = sla.generators.kregular_host(n, m, k, seed=21, planted=(kind == "forward"))
    v = v.astype(np.float64)
    v[7::1001] = 0.0
    v[11::1003] = 65535.0
    out = []
    for narrow in (1, 0):
        solver, z = getattr(sla, cls_name).new(n, m, n * k)
        solver.set_option("narrow_scan", narrow)
        solver.load_csr(n, m, rp, c, v.copy())
        solver.solve(z, maximize, None)
        assert solver.last_upload()[1] == 2
        assert solver.scan_value_bytes() == (2 if narrow else 8)
        assert solver.last_stats["wide_rounds"] >= 2                # both the zero-price and the gathering scan ran
        first = (z.person_to_object.copy(), z.object_to_person.copy(), solver.prices().copy(), dict(solver.last_stats))
        if narrow and kind == "khosla":
            assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v.copy(), maximize=maximize)
        solver.solve(z, maximize, None)                             # resident CSR again: same mirror, same results
        assert np.array_equal(z.person_to_object, first[0]) and np.array_equal(solver.prices(), first[2])
        out.append(first)
        if narrow:
            w = v + 0.25                                            # not u16: the next upload must drop the mirror
            solver.load_csr(n, m, rp, c, w.copy())
            solver.solve(z, maximize, None)
            assert solver.last_upload()[1] == 4 and solver.scan_value_bytes() == 8
            ref, zr = gpu_solve(sla, cls_name, n, m, rp, c, w.copy(), maximize=maximize, options={"narrow_upload": 0})
            assert np.array_equal(z.person_to_object, zr.person_to_object) and np.array_equal(solver.prices(), ref.prices())
    for x, y in zip(out[0][:3], out[1][:3]):
        assert np.array_equal(x, y)
    for key in ("num_unassigned", "nits", "nreductions", "rounds", "bids", "bid_arcs", "dropped", "values_negated", "eps"):
        assert out[0][3][key] == out[1][3][key], key


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_u16_index_type(sla, oracle, kind, cls_name):
    rng = np.random.default_rng(3)
    n, m, k = 300, 300 if kind == "forward" else 400, 6
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True)
    solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v, eps=1.0 / (m + 1), index_dtype=np.uint16)
    assert z.person_to_object.dtype == np.uint16 and z.object_to_person.dtype == np.uint16
    r = oracle.jacobi_model(kind, n, m, rp, c, v, eps=1.0 / (m + 1))
    assert np.array_equal(z.person_to_object, r["p2o"].astype(np.uint16))    # u32::MAX truncates to u16::MAX
    assert np.array_equal(z.object_to_person, r["o2p"].astype(np.uint16))
    if m > n:
        assert np.any(z.object_to_person == 0xFFFF)


def test_sign_handling_across_repeated_solves(sla, oracle):
    """solver.rs:207-216 is stateful: values are negated in place and the next solve sees the negated values."""
    rng = np.random.default_rng(11)
    n, m, k = 200, 300, 8
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=500)
    solver, z = sla.KhoslaSolver.new(n, m, n * k)
    solver.load_csr(n, m, rp, c, v)
    o = oracle.OracleSolver("khosla", n, m, n * k)
    o.load_csr(n, m, rp, c, v)
    for maximize in (False, False, True, False, True, True):
        solver.solve(z, maximize, 1.0 / (m + 1))
        o.solve(maximize=maximize, eps=1.0 / (m + 1))
        assert np.array_equal(solver.values(), o.values)
        assert solver.get_objective(z) == o.get_objective()
        assert z.num_unassigned == o.num_unassigned == 0
        assert solver.device_objective() == o.get_objective()      # device-side sum, exact on integer weights
        un, ok = solver.device_validate_matching()
        assert (un, ok) == (0, True)


def test_forward_params_and_eps_scaling(sla, oracle):
    """solve_with_params (symmetric.rs:217-332): start_eps, max_iterations, nreductions, optimal_soln_found, eps."""
    rng = np.random.default_rng(5)
    n, k = 500, 10
    rp, c, v = random_sparse_instance(rng, n, n, k, integer=True, lo=0, hi=1000)
    for kw in (dict(), dict(start_eps=50.0), dict(start_eps=1e-4), dict(max_iterations=7), dict(max_iterations=1),
               dict(start_eps=5.0, max_iterations=100)):
        solver, z = gpu_solve(sla, "ForwardAuctionSolver", n, n, rp, c, v, eps=1.0 / (n + 1), **(kw or {"start_eps": None}))
        o = oracle.OracleSolver("forward", n, n, n * k)
        o.load_csr(n, n, rp, c, v)
        o.solve(eps=1.0 / (n + 1), **kw)
        r = assert_equals_model(oracle, "forward", solver, z, n, n, rp, c, v, eps=1.0 / (n + 1), **kw)
        assert solver.optimal_soln_found == o.optimal_soln_found, kw
        if o.optimal_soln_found:      # a solve cut short by max_iterations stops mid-auction: only the model is comparable
            assert z.num_unassigned == o.num_unassigned == 0, kw
            assert solver.get_objective(z) == o.get_objective()
            assert solver.nreductions == o.nreductions and z.eps == o.eps
            tol = solver.get_toleration(float(np.max(np.abs(v))))
            assert solver.ecs_satisfied(z.person_to_object, 1.0 / (n + 1), tol)
            assert solver.device_ecs_satisfied(1.0 / (n + 1), tol)
        del r


def test_invalid_inputs_fail_loudly(sla):
    solver, z = sla.KhoslaSolver.new(4, 4, 16)
    solver.init(2, 3)
    solver.extend_from_values(0, [0, 7], [1.0, 2.0])          # column 7 >= num_cols (reference: debug_assert only)
    solver.extend_from_values(1, [1], [1.0])
    with pytest.raises(sla.SlaError) as e:
        solver.solve(z, False, None)
    assert e.value.code == 1 and "column" in e.value.message
    with pytest.raises(sla.SlaError):
        solver.set_option("no_such_option", 1)
    with pytest.raises(sla.SlaError):
        solver.set_option("tail_max", 4096)                  # the tail engine's smem queue holds 1024 bidders


@pytest.mark.parametrize("where", ["first", "middle", "last_odd_tail"])
@pytest.mark.parametrize("narrow", [1, 0])
def test_invalid_large_upload_fails_loudly_and_leaves_values_alone(sla, where, narrow):
    """The device-side validation of a large upload (csr_stats_kernel, 4 arcs per thread + a scalar tail): an out-of-range
    column anywhere in a 1.12 M-arc array is reported, the solve does not run, and the host `values` are back to (or
    still at) what the caller handed in -- the negation belongs to a solve that happens."""
    n, m, k = 70_001, 200_000, 16
    rp, c, v = sla.generators.kregular_host(n, m, k, seed=2)
    c = c.copy()
    pos = {"first": 0, "middle": c.size // 2 + 1, "last_odd_tail": c.size - 1}[where]
    c[pos] = m
    solver, z = sla.KhoslaSolver.new(n, m, n * k)
    solver.set_option("narrow_upload", narrow)
    solver.load_csr(n, m, rp, c, v)
    with pytest.raises(sla.SlaError) as e:
        solver.solve(z, False, None)
    assert e.value.code == 1 and "column" in e.value.message
    assert np.array_equal(solver.values(), v)                 # not negated: the solve never happened
    good, z2 = sla.KhoslaSolver.new(n, m, n * k)
    good.load_csr(n, m, rp, sla.generators.kregular_host(n, m, k, seed=2)[1], v)
    good.solve(z2, False, None)
    assert z2.num_unassigned == 0


def test_wide_first_round_with_most_persons_losing(sla, oracle):
    """Plain Khosla with 641 .. 1,024 persons runs its first round on the grid-wide kernels inside a one-super-round
    graph; when more than half of the persons lose that round (everybody prefers the same few objects) the queue is still
    too long for the tail engine and continuation graphs have to take over -- same bits as the model either way."""
    rng = np.random.default_rng(21)
    n, m, k = 900, 2000, 8
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=50)
    v = v.reshape(n, k)
    c2 = c.reshape(n, k).copy()
    c2[:, 0] = np.arange(n) % 3                       # three hot objects in everybody's row ...
    c2.sort(axis=1)
    for i in range(n):                                # (keep rows duplicate-free and sorted)
        while len(set(c2[i])) < k:
            c2[i] = np.sort(np.unique(np.concatenate([np.unique(c2[i]), rng.choice(m, size=k, replace=False)]))[:k])
    hot = c2 < 3
    v[hot] = 1000.0                                   # ... and far more valuable than anything else (maximize)
    c, v = c2.reshape(-1).astype(np.uint32), v.reshape(-1)
    solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, v, maximize=True)
    assert solver.last_stats["wide_rounds"] >= 2 and solver.last_stats["graph_launches"] >= 2
    assert_equals_model(oracle, "khosla", solver, z, n, m, rp, c, v, maximize=True)


# ---- BASELINE.json configurations -----------------------------------------------------------------------------------
def test_cfg1_khosla_1000x10000_k32(sla, oracle):
    from sparse_linear_assignment_b200 import generators as G
    n, m, k, _ = G.CONFIGS["cfg1"]
    rp, c, v = G.kregular_host(n, m, k, seed=1, value_dist="beta33")          # (1a) floor(700*Beta(3,3)+300), benchmark.rs:60,73
    assert v.min() >= 300 and v.max() < 1000 and abs(v.mean() - 650) < 5 and abs(v.std() - 132.3) < 4
    solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, v)                 # integer costs, eps=None
    o = oracle.OracleSolver("khosla", n, m, n * k)
    o.load_csr(n, m, rp, c, v)
    o.solve()
    assert z.num_unassigned == o.num_unassigned == 0
    assert solver.get_objective(z) == o.get_objective()                        # eps = 1/M < 1/N: exact regime
    assert_equals_model(oracle, "khosla", solver, z, n, m, rp, c, v)
    rng = np.random.default_rng(1)
    vf = rng.uniform(0, 10, size=n * k)                                       # (1b) non-integer weights
    solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, vf)
    o.load_csr(n, m, rp, c, vf)
    o.solve()
    assert z.num_unassigned == o.num_unassigned == 0
    assert abs(solver.get_objective(z) - o.get_objective()) <= n * z.eps      # n * eps_final


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_device_generator_equals_host_generator(sla, oracle, kind, cls_name):
    from sparse_linear_assignment_b200 import generators as G
    n, m, k = 2048, (2048 if kind == "forward" else 6000), 16
    planted = kind == "forward"
    rp, c, v = G.kregular_host(n, m, k, seed=5, planted=planted)
    assert np.all(np.diff(c.reshape(n, k), axis=1) > 0) and c.max() < m and v.min() >= 300 and v.max() < 1000
    host, zh = gpu_solve(sla, cls_name, n, m, rp, c, v, eps=1.0 / (m + 1))
    dev, zd = getattr(sla, cls_name).new(n, m, n * k)
    G.kregular_device(dev, n, m, k, seed=5, planted=planted)
    dev.solve(zd, False, 1.0 / (m + 1))
    assert np.array_equal(zh.person_to_object, zd.person_to_object)
    assert np.array_equal(host.prices(), dev.prices())
    assert host.last_stats["bid_arcs"] == dev.last_stats["bid_arcs"]
    assert zd.num_unassigned == 0
    assert dev.device_objective() == host.get_objective(zh)


def test_cfg2_forward_20000x20000_k64(sla, oracle):
    from sparse_linear_assignment_b200 import generators as G
    n, m, k, planted = G.CONFIGS["cfg2"]
    rp, c, v = G.kregular_host(n, m, k, seed=1, planted=planted)
    eps = 1.0 / (n + 1)                                                        # strict < 1/n: bit-exact parity run
    solver, z = gpu_solve(sla, "ForwardAuctionSolver", n, m, rp, c, v, eps=eps)
    o = oracle.OracleSolver("forward", n, m, n * k)
    o.load_csr(n, m, rp, c, v)
    o.solve(eps=eps)
    assert z.num_unassigned == o.num_unassigned == 0
    assert solver.optimal_soln_found and o.optimal_soln_found
    assert solver.get_objective(z) == o.get_objective()
    check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, 0)
    assert_equals_model(oracle, "forward", solver, z, n, m, rp, c, v, eps=eps)


def test_cfg3_khosla_1Mx4M_k16(sla, oracle):
    from sparse_linear_assignment_b200 import generators as G
    n, m, k, _ = G.CONFIGS["cfg3"]
    rp, c, v = G.kregular_host(n, m, k, seed=1)
    solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, v)                 # eps=None -> 1/M < 1/N
    o = oracle.OracleSolver("khosla", n, m, n * k)
    o.load_csr(n, m, rp, c, v)
    o.solve()
    assert z.num_unassigned == o.num_unassigned == 0
    assert solver.get_objective(z) == o.get_objective()
    check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, 0)
    assert solver.device_objective() == o.get_objective()
    assert solver.device_validate_matching() == (0, True)
    assert_equals_model(oracle, "khosla", solver, z, n, m, rp, c, v)
    # the same instance generated in HBM, solved without any host copy
    dev, zd = sla.KhoslaSolver.new(n, m, n * k)
    G.kregular_device(dev, n, m, k, seed=1)
    st = dev.solve_resident(False, None)
    assert st["bid_arcs"] == solver.last_stats["bid_arcs"] and st["num_unassigned"] == 0
    assert dev.device_objective() == o.get_objective()


def _symmetric_instance(n, mean_degree, seed, planted, lo, hi):
    rng = np.random.default_rng(seed)
    perm = rng.permutation(n)
    rows = []
    for i in range(n):
        cc = rng.choice(n, size=max(int(rng.binomial(n, mean_degree / n)), 1), replace=False)
        rows.append(np.unique(np.append(cc, perm[i]) if planted else cc))
    rp = np.zeros(n + 1, dtype=np.uint32)
    rp[1:] = np.cumsum([len(r) for r in rows])
    c = np.concatenate(rows).astype(np.uint32)
    return rp, c, rng.uniform(lo, hi, size=c.size)


def test_khosla_schedule_is_engine_independent(sla, oracle):
    """The Khosla eps-schedule on a square instance through every engine split (wide only / single-CTA only / with and
    without the shared-memory mirrors / host loop): identical bits, equal to the model."""
    rng = np.random.default_rng(11)
    n, k = 2500, 24
    rp, c, v = random_sparse_instance(rng, n, n, k, integer=True, lo=0, hi=500)      # plants a perfect matching
    eps = 1.0 / (n + 1)
    ref = oracle.jacobi_model("khosla", n, n, rp, c, v, eps=eps)
    assert ref["stats"]["num_unassigned"] == 0 and ref["stats"]["nreductions"] >= 5
    for opt in (dict(), dict(graph=0), dict(tail_max=0), dict(tail_max=40, super_rounds=3),
                dict(smem_prices=0), dict(smem_owners=0), dict(regular=0, tail_max=700)):
        solver, z = gpu_solve(sla, "KhoslaSolver", n, n, rp, c, v, eps=eps, options=opt)
        assert np.array_equal(z.person_to_object, ref["p2o"]), opt
        assert np.array_equal(solver.prices(), ref["prices"]), opt
        for key in ("num_unassigned", "nits", "nreductions", "rounds", "bids", "bid_arcs"):
            assert solver.last_stats[key] == ref["stats"][key], (opt, key)


def test_khosla_eps_schedule_on_square_instances(sla, oracle):
    """Khosla rounds on square instances run under an eps-schedule that ends at the caller's eps: same
    eps-CS guarantee as the reference's fixed-eps loop, an order of magnitude fewer rounds.  A phase that drops anybody
    at the price threshold makes the solve start over with the plain rounds, so instances without a perfect matching
    behave exactly as before.  Both paths equal the CPU model bit for bit; objectives are judged against the oracle."""
    n = 1200
    rp, c, v = _symmetric_instance(n, 12, 3, True, 500.0, 1000.0)           # the reference's symmetric bench shape
    o = oracle.OracleSolver("khosla", n, n, len(c))
    o.load_csr(n, n, rp, c, v)
    o.solve()
    solver, z = gpu_solve(sla, "KhoslaSolver", n, n, rp, c, v)
    assert z.num_unassigned == o.num_unassigned == 0 and z.eps == o.eps
    assert abs(solver.get_objective(z) - o.get_objective()) <= n * z.eps    # n * eps_final (non-integer weights)
    assert solver.last_stats["nreductions"] >= 5 and solver.last_stats["restarts"] == 0
    assert_equals_model(oracle, "khosla", solver, z, n, n, rp, c, v)
    scaled_rounds = solver.last_stats["rounds"]
    plain, zp = gpu_solve(sla, "KhoslaSolver", n, n, rp, c, v, options=dict(khosla_scaling=0))
    assert zp.num_unassigned == 0 and plain.last_stats["nreductions"] == 0 and plain.last_stats["restarts"] == 0
    assert abs(plain.get_objective(zp) - o.get_objective()) <= n * zp.eps
    assert_equals_model(oracle, "khosla", plain, zp, n, n, rp, c, v, khosla_scaling=False)
    assert scaled_rounds * 3 < plain.last_stats["rounds"]
    # integer weights, eps < 1/n: bit-exact objective either way
    vi = np.floor(v)
    o.load_csr(n, n, rp, c, vi)
    o.solve(eps=1.0 / (n + 1))
    solver, z = gpu_solve(sla, "KhoslaSolver", n, n, rp, c, vi, eps=1.0 / (n + 1))
    assert z.num_unassigned == 0 and solver.get_objective(z) == o.get_objective()
    assert_equals_model(oracle, "khosla", solver, z, n, n, rp, c, vi, eps=1.0 / (n + 1))
    # no perfect matching: the schedule is abandoned, results equal the plain rounds' and the oracle's num_unassigned
    infeasible = 0
    for nn, seed in ((9, 5), (60, 6), (150, 7), (40, 8)):
        rp, c, v = _symmetric_instance(nn, 3, seed, False, 0.0, 10.0)
        o = oracle.OracleSolver("khosla", nn, nn, len(c))
        o.load_csr(nn, nn, rp, c, v)
        o.solve()
        solver, z = gpu_solve(sla, "KhoslaSolver", nn, nn, rp, c, v)
        assert z.num_unassigned == o.num_unassigned
        check_matching(nn, nn, rp, c, z.person_to_object, z.object_to_person, o.num_unassigned)
        r = assert_equals_model(oracle, "khosla", solver, z, nn, nn, rp, c, v)
        if o.num_unassigned:
            infeasible += 1
            assert solver.last_stats["restarts"] == 1              # sla_stats reports the abandoned schedule
            plain = oracle.jacobi_model("khosla", nn, nn, rp, c, v, khosla_scaling=False)
            assert np.array_equal(r["p2o"], plain["p2o"]) and np.array_equal(r["prices"], plain["prices"])
    assert infeasible >= 2


# ---- edge cases of the reference's input space ------------------------------------------------------------------------
@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_edge_cases_match_model_and_oracle(sla, oracle, kind, cls_name):
    rng = np.random.default_rng(77)
    cases = {}
    # one person, one object
    cases["1x1"] = (1, 1, np.array([0, 1], dtype=np.uint32), np.array([0], dtype=np.uint32), np.array([5.0]))
    # all weights equal: every comparison is a tie (lowest row position / lowest person id must win)
    n = 40
    cases["all_ties"] = (n, n, np.arange(0, n * n + 1, n, dtype=np.uint32), np.tile(np.arange(n, dtype=np.uint32), n),
                         np.full(n * n, 7.0))
    # duplicate column indices inside a row (the builder does not forbid them, solver.rs:41-101)
    rp = np.arange(0, 30 * 6 + 1, 6, dtype=np.uint32)
    c = rng.integers(0, 45, size=30 * 6).astype(np.uint32)
    c[::6] = np.arange(30, dtype=np.uint32)                     # keep a perfect matching available
    cases["duplicate_columns"] = (30, 45, rp, c, rng.integers(1, 50, size=30 * 6).astype(np.float64))
    # negative weights, maximise
    r2, c2, v2 = random_sparse_instance(rng, 64, 80, 9, integer=True, lo=-500, hi=-1)
    cases["negative_maximize"] = (64, 80, r2, c2, v2)
    # long rows: 200 arcs per person (several 128-bit chunks per lane, 32 lanes per row)
    r3, c3, v3 = random_sparse_instance(rng, 50, 400, 200, integer=True, lo=0, hi=10_000)
    cases["long_rows"] = (50, 400, r3, c3, v3)
    # dense 64 x 64
    cases["dense64"] = (64, 64, np.arange(0, 64 * 64 + 1, 64, dtype=np.uint32), np.tile(np.arange(64, dtype=np.uint32), 64),
                        rng.integers(0, 1000, size=64 * 64).astype(np.float64))
    # huge magnitudes next to tiny ones (f64 bids far from the 32-bit range)
    r4, c4, v4 = random_sparse_instance(rng, 33, 50, 5, integer=True, lo=1, hi=9)
    cases["wide_dynamic_range"] = (33, 50, r4, c4, v4 * np.where(rng.random(v4.size) < 0.5, 1e9, 1.0))
    for name, (n, m, rp, c, v) in cases.items():
        maximize = name == "negative_maximize"
        eps = 1.0 / (m + 1)
        solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v, maximize=maximize, eps=eps)
        assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v, maximize=maximize, eps=eps)
        o = oracle.OracleSolver(kind, n, m, len(c))
        o.load_csr(n, m, rp, c, v)
        o.solve(maximize=maximize, eps=eps)
        assert z.num_unassigned == o.num_unassigned, name
        if z.num_unassigned == 0 and name != "wide_dynamic_range":
            assert solver.get_objective(z) == o.get_objective(), name
        if z.num_unassigned == 0 and name == "wide_dynamic_range":
            assert abs(solver.get_objective(z) - o.get_objective()) <= n * eps + 1e-6, name
        if name != "duplicate_columns":
            check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, z.num_unassigned)


# ---- randomized sweep: shapes, degrees, float / integer weights, maximise / minimise, explicit / default eps -----------
def test_randomized_sweep_matches_model(sla, oracle):
    rng = np.random.default_rng(2024)
    for trial in range(60):
        kind, cls_name = SOLVERS[trial % 2]
        n = int(rng.integers(1, 700))
        m = n if (kind == "forward" and trial % 3) else n + int(rng.integers(0, 500))
        kmax = int(min(m, rng.integers(1, 70)))
        rp, c, v = ragged_instance(rng, n, m, 1 if trial % 7 == 0 and kind == "khosla" else min(2, kmax), max(kmax, min(2, m)),
                                   integer=bool(trial % 4))
        if trial % 5 == 0:
            v = -v - 1.0                                         # all-negative weights
        maximize = bool(trial % 3 == 0)
        eps = None if trial % 2 else float(rng.uniform(1e-4, 0.5))
        solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v, maximize=maximize, eps=eps)
        assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v, maximize=maximize, eps=eps)
        check_matching(n, m, rp, c, z.person_to_object, z.object_to_person, z.num_unassigned)
        assert_reference_parity(oracle, kind, solver, z, n, m, rp, c, v, maximize=maximize, eps=eps, integer=bool(trial % 4))


def test_randomized_square_sweep_matches_model(sla, oracle):
    """Square instances of every size class of the engines (<= 24 rows: slot-stable rounds from the start; <= 1024:
    single-CTA engine only; larger: wide rounds first), rows from 1 to 200 arcs (register rows of 32 / 64 / 128 arcs
    and the streaming fallback), planted (perfect matching: Khosla rounds under the eps-schedule) and unplanted
    (possibly none: the schedule is abandoned) -- both solvers against the CPU model, bit for bit."""
    rng = np.random.default_rng(77)
    for trial in range(48):
        kind, cls_name = SOLVERS[trial % 2]
        n = int((rng.integers(1, 25), rng.integers(25, 1025), rng.integers(1025, 3000))[trial % 3])
        kmax = int(min(n, (4, 20, 70, 140, 200)[trial % 5]))
        integer = bool(trial % 4)
        if trial % 6 == 5 and n <= 1024:
            nn = min(n, 60)                                      # unplanted and sparse: often without a perfect matching
            rp, c, v = _symmetric_instance(nn, 3, 100 + trial, False, 0.0, 10.0)
            n = nn
            if kind == "forward":
                continue                                         # the Forward solver runs to max_iterations there
        else:
            rp, c, v = ragged_instance(rng, n, n, 1 if n == 1 else min(2, kmax), max(kmax, min(2, n)), integer=integer)
        maximize = bool(trial % 3 == 0)
        eps = None if trial % 2 else float(rng.uniform(1e-4, 0.5))
        solver, z = gpu_solve(sla, cls_name, n, n, rp, c, v, maximize=maximize, eps=eps)
        assert_equals_model(oracle, kind, solver, z, n, n, rp, c, v, maximize=maximize, eps=eps)
        check_matching(n, n, rp, c, z.person_to_object, z.object_to_person, z.num_unassigned)
        assert_reference_parity(oracle, kind, solver, z, n, n, rp, c, v, maximize=maximize, eps=eps,
                                integer=bool(np.all(v == np.floor(v))))


# ---- Khosla on square instances around the feasibility boundary: num_unassigned is the reference's ---------------------
def test_square_khosla_feasibility_boundary_equals_reference(sla, oracle):
    """KhoslaSolver drops a person for good once the price of its best object exceeds the threshold
    (ksparse.rs:181, 218-220); on a square instance that is what decides `num_unassigned`.  The device runs such
    instances under an eps-schedule and falls back to plain rounds when a phase drops anybody.  Property,
    on sparse random square graphs of mean degree 2 .. 4 -- right at the threshold where a perfect matching stops
    existing -- with the schedule ON: the device's num_unassigned equals the reference's, for 200 unplanted instances
    (n = 10 .. 200; three quarters of them without a perfect matching) and 40 planted ones (n = 200 .. 2,000; a perfect
    matching exists, so nobody may be dropped), integer and real weights; the matching is valid; with everybody assigned
    the objective meets the north star's bar."""
    solver, z = sla.KhoslaSolver.new(2048, 2048, 2048 * 12)
    infeasible = feasible = 0
    for seed in range(240):
        rng = np.random.default_rng(seed)
        planted = seed >= 200
        n = int(rng.integers(200, 2001)) if planted else int(rng.integers(10, 201))
        integer = seed % 2 == 0
        rp, c, v = _symmetric_instance(n, float(rng.uniform(2.0, 4.0)), 5000 + seed, planted, 0.0, 50.0 if integer else 10.0)
        if integer:
            v = np.floor(v)
        solver.load_csr(n, n, rp, c, v.copy())
        solver.solve(z, False, None)
        o = assert_reference_parity(oracle, "khosla", solver, z, n, n, rp, c, v, integer=integer)
        check_matching(n, n, rp, c, z.person_to_object, z.object_to_person, o.num_unassigned)
        if planted:
            assert z.num_unassigned == 0
        if o.num_unassigned:
            assert solver.last_stats["restarts"] == 1              # nobody ends unassigned under the schedule itself
        infeasible += int(o.num_unassigned > 0)
        feasible += int(o.num_unassigned == 0)
    assert infeasible >= 100 and feasible >= 50


@pytest.mark.parametrize("kind,cls_name", SOLVERS)
def test_square_real_weight_fixtures_against_reference(sla, oracle, kind, cls_name):
    """Square instances with real-valued weights (the reference's own fixture family, solver.rs:261-292, and its
    symmetric bench shape, benches/benchmark.rs:16-47) against the reference oracle: Khosla runs them under the
    eps-schedule, so equality with the reference's prices is not promised -- eps-CS at the caller's eps is, and with it
    the objective within n * eps."""
    for n, k, seed in ((60, 6, 1), (300, 10, 2), (1200, 12, 3)):
        rp, c, v = oracle.fixture_ksparse(n, n, k, 10.0, val_seed=seed, filter_seed=seed + 100)
        rp2, c2, v2 = _symmetric_instance(n, k, 40 + seed, True, 500.0, 1000.0)
        for (a, b, w) in ((rp, c, v), (rp2, c2, v2)):
            solver, z = gpu_solve(sla, cls_name, n, n, a, b, w.copy())
            o = assert_reference_parity(oracle, kind, solver, z, n, n, a, b, w)
            check_matching(n, n, a, b, z.person_to_object, z.object_to_person, z.num_unassigned)
            if z.num_unassigned == 0:
                # eps-CS as solver.rs:154-189 tests it; Khosla's update rule (ksparse.rs:222-227) guarantees it in
                # exact arithmetic only, so its check gets a few ulps of the weight range on top of the toleration
                tol = solver.get_toleration(float(np.max(np.abs(w))))
                e = 1.0 / n if kind == "forward" else z.eps + 64 * tol
                assert solver.device_ecs_satisfied(e, tol)
            del o


def test_failed_solve_does_not_poison_the_next_one(sla, oracle):
    """A solve that fails after the upload has pre-applied the sign normalisation (here: the wall-clock guard) must leave
    host and device agreeing about the sign, so that the next solve on the unchanged CSR just works."""
    rng = np.random.default_rng(4)
    n, k = 3000, 16
    rp, c, v = random_sparse_instance(rng, n, n, k, integer=True, lo=1, hi=1000)
    solver, z = sla.ForwardAuctionSolver.new(n, n, n * k)
    solver.load_csr(n, n, rp, c, v.copy())
    solver.set_option("super_rounds", 1)
    solver.set_option("timeout_s", 0)
    with pytest.raises(sla.SlaError) as e:
        solver.solve(z, False, 1.0 / (n + 1))                # needs far more than one graph launch
    assert "timeout" in e.value.message or "guard" in e.value.message
    assert np.array_equal(solver.values(), -v)               # the normalisation happened (solver.rs:214-216) ...
    solver.set_option("timeout_s", 900)
    solver.set_option("super_rounds", 6)
    solver.solve(z, False, 1.0 / (n + 1))                    # ... and the next solve agrees with it
    assert np.array_equal(solver.values(), -v)
    o = oracle.OracleSolver("forward", n, n, n * k)
    o.load_csr(n, n, rp, c, v)
    o.solve(eps=1.0 / (n + 1))
    assert z.num_unassigned == 0 and solver.get_objective(z) == o.get_objective()


def test_bound_pruned_gather_changes_nothing(sla, oracle):
    """Rounds with >= 32 Ki bidders gather prices only for arcs that can still matter (value >= a proven bound on the
    second-best profit); option prune_gather = 0 gathers everything.  Identical bits either way, equal to the model: u16
    and f64 values, both signs, a first round that gathers (zero_price_skip = 0), K = 8 .. 64."""
    for n, m, k, seed in ((70_000, 200_000, 16, 1), (40_000, 90_000, 8, 2), (36_000, 120_000, 40, 3), (34_000, 34_000, 64, 4)):
        rp, c, v = sla.generators.kregular_host(n, m, k, seed=seed, planted=(n == m))
        for vals in (v, v + 0.25):                           # u16 mirror / f64 values (f32 on the wire)
            for maximize in (False, True):
                out = []
                for prune in (1, 0):
                    solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, vals.copy(), maximize=maximize,
                                          options=dict(prune_gather=prune, zero_price_skip=0, khosla_scaling=0))
                    assert solver.last_stats["wide_rounds"] >= 2
                    out.append((z.person_to_object.copy(), z.object_to_person.copy(), solver.prices().copy(),
                                {q: solver.last_stats[q] for q in ("rounds", "bids", "bid_arcs", "num_unassigned", "dropped")}))
                for x, y in zip(out[0][:3], out[1][:3]):
                    assert np.array_equal(x, y)
                assert out[0][3] == out[1][3]
        solver, z = gpu_solve(sla, "KhoslaSolver", n, m, rp, c, v.copy(), options=dict(zero_price_skip=0, khosla_scaling=0))
        assert_equals_model(oracle, "khosla", solver, z, n, m, rp, c, v.copy(), khosla_scaling=False)


def test_cluster_engine_changes_nothing(sla, oracle):
    """Instances whose prices do not fit one CTA's shared memory run their queues of 25 .. 8,192 bidders in the cluster
    engine (one thread-block cluster, queue split over its CTAs' shared memory).  Option cluster_engine = 0 leaves those
    rounds to the grid-wide pair and the single-CTA engine.  Identical bits either way and equal to the model: uniform
    and ragged rows, both solvers, both signs, eps phases with resets (square Khosla, Forward), the engine as the first
    one of a solve (small N, smem_prices = 0), other hand-over points."""
    rng = np.random.default_rng(11)
    cases = []
    rp, c, v = sla.generators.kregular_host(150_000, 600_000, 16, seed=5)
    cases.append(("khosla", "KhoslaSolver", 150_000, 600_000, rp, c, v, False, None, dict(khosla_scaling=0), {}))
    cases.append(("khosla", "KhoslaSolver", 150_000, 600_000, rp, c, v + 0.5, True, None, dict(khosla_scaling=0), {}))
    rp, c, v = ragged_instance(rng, 6_000, 40_000, 2, 24, True)
    cases.append(("khosla", "KhoslaSolver", 6_000, 40_000, rp, c, v, False, 1.0 / 40_001, dict(khosla_scaling=0), {}))
    rp, c, v = sla.generators.kregular_host(36_000, 36_000, 16, seed=6, planted=True)
    cases.append(("khosla", "KhoslaSolver", 36_000, 36_000, rp, c, v, False, 1.0 / 36_001, {}, {}))          # eps-schedule
    cases.append(("forward", "ForwardAuctionSolver", 36_000, 36_000, rp, c, v, True, 1.0 / 36_001, {}, {}))
    rp, c, v = random_sparse_instance(rng, 3000, 5000, 16, integer=True, lo=0, hi=200)
    cases.append(("khosla", "KhoslaSolver", 3000, 5000, rp, c, v, False, 1.0 / 5001, dict(smem_prices=0, khosla_scaling=0),
                  dict(khosla_scaling=False)))
    rp, c, v = random_sparse_instance(rng, 3000, 3000, 16, integer=True, lo=0, hi=200)
    cases.append(("forward", "ForwardAuctionSolver", 3000, 3000, rp, c, v, False, 1.0 / 3001, dict(smem_prices=0), {}))
    for kind, cls_name, n, m, rp, c, v, maximize, eps, opts, model_kw in cases:
        if "khosla_scaling" in opts and kind == "khosla":
            model_kw = dict(khosla_scaling=False)
        out = []
        for extra in (dict(cluster_engine=1), dict(cluster_engine=0), dict(cluster_engine=1, cluster_handover=1),
                      dict(cluster_engine=1, cluster_handover=300)):
            solver, z = gpu_solve(sla, cls_name, n, m, rp, c, v.copy(), maximize=maximize, eps=eps,
                                  options=dict(opts, timeout_s=600, **extra))
            st = solver.last_stats
            assert st["wide_rounds"] + st["tail_rounds"] == st["rounds"] and st["cluster_rounds"] <= st["tail_rounds"]
            assert (st["cluster_rounds"] > 0) == bool(extra["cluster_engine"]), (kind, n, m, extra, st)
            out.append((z.person_to_object.copy(), z.object_to_person.copy(), solver.prices().copy(),
                        {q: st[q] for q in ("rounds", "bids", "bid_arcs", "num_unassigned", "dropped", "nits", "nreductions",
                                            "optimal_soln_found", "eps")}))
            if extra == dict(cluster_engine=1):
                assert_equals_model(oracle, kind, solver, z, n, m, rp, c, v.copy(), maximize=maximize, eps=eps, **model_kw)
        for other in out[1:]:
            for x, y in zip(out[0][:3], other[:3]):
                assert np.array_equal(x, y)
            assert out[0][3] == other[3]


def test_resident_repeat_solves_learn_the_graph_shape(sla, oracle):
    """Repeated plain-Khosla solves of a resident CSR capture their first graph as (wide rounds of the previous solve) x
    (bid, assign) + one tail launch; results and counters stay identical, launches go down, and a solve whose round count
    differs (another eps) is finished by continuation graphs."""
    n, m, k = 300_000, 1_200_000, 16
    solver, z = sla.KhoslaSolver.new(n, m, n * k)
    sla.generators.kregular_device(solver, n, m, k, seed=3)
    a = solver.solve_resident(False, None)
    first = (a["rounds"], a["bids"], a["bid_arcs"], a["num_unassigned"])
    obj = solver.device_objective()
    b = solver.solve_resident(False, None)
    c2 = solver.solve_resident(False, None)
    for st in (b, c2):
        assert (st["rounds"], st["bids"], st["bid_arcs"], st["num_unassigned"]) == first
        assert st["kernel_launches"] < a["kernel_launches"] and st["graph_launches"] == 1
    assert solver.device_objective() == obj
    d = solver.solve_resident(False, 5.0)                    # coarser eps: fewer rounds than the learned shape expects
    e = solver.solve_resident(False, None)                   # and back
    assert (e["rounds"], e["bids"], e["bid_arcs"], e["num_unassigned"]) == first and d["num_unassigned"] == 0
    assert solver.device_objective() == obj
    solver.set_option("learn_shape", 0)
    f = solver.solve_resident(False, None)
    assert (f["rounds"], f["bids"], f["bid_arcs"]) == first[:3] and f["kernel_launches"] == a["kernel_launches"]


def test_cfg5_khosla_16Mx64M_k16(sla, oracle):
    """BASELINE.json config 5 at full size on one GPU, from host memory, against the reference oracle: same
    num_unassigned, bit-exact objective (integer costs, eps = 1/M < 1/N), valid matching (device-side validator)."""
    from sparse_linear_assignment_b200 import generators as G
    n, m, k, _ = G.CONFIGS["cfg5"]
    solver, z = sla.KhoslaSolver.new(n, m, n * k)
    solver.init(n, m)
    solver._i_starts_stops.resize(n + 1, 0)                   # generate straight into the solver's (page-locked) storage
    solver._j_counts.resize(n, k)
    solver._column_indices.resize(n * k, 0)
    solver._values.resize(n * k, 0.0)
    G.kregular_host(n, m, k, seed=1, out=(solver._i_starts_stops.view, solver._column_indices.view, solver._values.view))
    o = oracle.OracleSolver("khosla", n, m, n * k)
    o.load_csr(n, m, solver.i_starts_stops(), solver.column_indices(), solver.values())
    solver.solve(z, False, None)
    o.solve()
    assert z.num_unassigned == o.num_unassigned == 0
    assert solver.device_validate_matching() == (0, True)
    assert solver.device_objective() == o.get_objective()      # exact: integer costs
    assert np.array_equal(solver.values(), o.values)          # both negated in place
    p2o = z.person_to_object
    assert np.array_equal(z.object_to_person[p2o], np.arange(n, dtype=np.uint32))
    # the same instance generated in HBM: identical work counters
    dev, zd = sla.KhoslaSolver.new(n, m, n * k)
    G.kregular_device(dev, n, m, k, seed=1)
    st = dev.solve_resident(False, None)
    assert st["bid_arcs"] == solver.last_stats["bid_arcs"] and st["num_unassigned"] == 0
    assert dev.device_objective() == o.get_objective()
