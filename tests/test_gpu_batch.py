"""GPU: the one-CTA-per-instance batch engine (BASELINE.json config 4) against the CPU model and the oracle."""
import numpy as np
import pytest

from helpers import check_matching, objective_of, random_sparse_instance

pytestmark = pytest.mark.gpu


def check_batch(sla, oracle, kind, instances, res, eps, **kw):
    b = res["_solver"]
    for idx, (n, m, rp, c, v) in enumerate(instances):
        rs, cs = b.instance_slices(idx)
        p2o, o2p, prices, st = res["p2o"][rs], res["o2p"][cs], res["prices"][cs], res["stats"][idx]
        model = oracle.jacobi_model(kind, n, m, rp, c, v, eps=eps, **kw)
        assert np.array_equal(p2o, model["p2o"]), idx
        assert np.array_equal(o2p, model["o2p"]), idx
        assert np.array_equal(prices, model["prices"]), idx
        for key in ("num_unassigned", "nits", "nreductions", "optimal_soln_found", "rounds", "bids", "bid_arcs", "dropped",
                    "values_negated", "restarts"):
            assert st[key] == model["stats"][key], (idx, key)
        assert st["eps"] == model["stats"]["eps"]
        # the Khosla eps-schedule only exists on square instances; abandoning it is reported, and whoever ends with
        # unassigned persons under the schedule has abandoned it (the plain rounds define the drops)
        assert st["restarts"] in (0, 1) and (st["restarts"] == 0 or (kind == "khosla" and n == m)), idx
        if kind == "khosla" and n == m and st["num_unassigned"] and not kw:
            assert st["restarts"] == 1, idx
        check_matching(n, m, rp, c, p2o, o2p, st["num_unassigned"])
        if not kw:
            o = oracle.OracleSolver(kind, n, m, len(c))
            o.load_csr(n, m, rp, c, v)
            o.solve(eps=eps)
            assert st["num_unassigned"] == o.num_unassigned == 0
            assert objective_of(rp, c, v, p2o) == o.get_objective()       # integer weights, eps < 1/n: bit-exact


@pytest.mark.parametrize("kind", ["khosla", "forward"])
def test_ragged_batch_of_mixed_sizes(sla, oracle, kind):
    rng = np.random.default_rng(21)
    instances = []
    for _ in range(40):
        n = int(rng.integers(3, 120))
        m = n if kind == "forward" and rng.random() < 0.7 else n + int(rng.integers(0, 60))
        k = int(rng.integers(2, min(m, 24) + 1))
        rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=300)
        instances.append((n, m, rp, c, v))
    eps = 1.0 / 400
    b = sla.BatchSolver(kind)
    b.upload(instances)
    res = b.solve(eps=eps)
    res["_solver"] = b
    check_batch(sla, oracle, kind, instances, res, eps)
    assert res["total"]["bid_arcs"] == sum(s["bid_arcs"] for s in res["stats"])
    assert res["total"]["restarts"] == sum(s["restarts"] for s in res["stats"])
    if kind == "forward":
        cut = b.solve(eps=eps, max_iterations=5)
        cut["_solver"] = b
        check_batch(sla, oracle, kind, instances, cut, eps, max_iterations=5)


@pytest.mark.parametrize("kind", ["forward", "khosla"])
def test_cfg4_shaped_batch_generated_on_device(sla, oracle, kind):
    from sparse_linear_assignment_b200 import generators as G
    n = m = 512
    k, count, first = 32, 24, 1000
    b = sla.BatchSolver(kind)
    b.generate_device(count, first, n, m, k, seed=0, planted=True)
    eps = 1.0 / (n + 1)
    res = b.solve(eps=eps)
    res["_solver"] = b
    instances = []
    for i in range(count):
        rp, c, v = G.kregular_host(n, m, k, seed=first + i, planted=True)
        instances.append((n, m, rp, c, v))
    check_batch(sla, oracle, kind, instances, res, eps)
    assert res["total"]["num_unassigned"] == 0


def test_cfg4_full_batch_8192_instances(sla, oracle):
    """BASELINE.json config 4 at its full count: 8,192 independent 512 x 512 k=32 instances (Forward, eps-scaled)
    generated in HBM and solved by one launch.  Every instance against the CPU model bit for bit (assignment, prices,
    counters); 256 of them against the reference oracle (num_unassigned, bit-exact objective: integer costs,
    eps = 1/(n+1) < 1/n)."""
    from concurrent.futures import ThreadPoolExecutor
    from sparse_linear_assignment_b200 import generators as G
    n = m = 512
    k, count = 32, 8192
    eps = 1.0 / (n + 1)
    b = sla.BatchSolver("forward")
    b.generate_device(count, 0, n, m, k, seed=0, planted=True)
    res = b.solve(eps=eps)
    assert res["total"]["num_unassigned"] == 0
    oracle.lib()

    def check(i):
        rp, c, v = G.kregular_host(n, m, k, seed=i, planted=True, threads=1)
        rs, cs = b.instance_slices(i)
        model = oracle.jacobi_model("forward", n, m, rp, c, v, eps=eps)
        ok = (np.array_equal(res["p2o"][rs], model["p2o"]) and np.array_equal(res["o2p"][cs], model["o2p"]) and
              np.array_equal(res["prices"][cs], model["prices"]))
        st = res["stats"][i]
        ok = ok and all(st[q] == model["stats"][q] for q in ("num_unassigned", "nits", "nreductions", "optimal_soln_found",
                                                             "rounds", "bids", "bid_arcs"))
        if ok and i % 32 == 0:
            o = oracle.OracleSolver("forward", n, m, n * k)
            o.load_csr(n, m, rp, c, v)
            o.solve(eps=eps)
            ok = o.num_unassigned == 0 and objective_of(rp, c, v, res["p2o"][rs]) == o.get_objective()
        return ok

    # jacobi_model keeps its option in a global: set it once, the worker threads then only read it
    oracle.jacobi_model("forward", 1, 1, np.array([0, 1], dtype=np.uint32), np.zeros(1, dtype=np.uint32), np.ones(1))
    with ThreadPoolExecutor(max_workers=16) as ex:
        bad = [i for i, ok in enumerate(ex.map(check, range(count))) if not ok]
    assert not bad, bad[:10]


def test_batch_errors(sla):
    b = sla.BatchSolver("khosla")
    with pytest.raises(sla.SlaError):
        b.solve()                                                   # nothing uploaded
    rp = np.array([0, 2, 4], dtype=np.uint32)
    with pytest.raises(sla.SlaError):
        b.upload([(2, 1, rp, np.zeros(4, dtype=np.uint32), np.ones(4))])   # rows > cols (solver.rs:192)
