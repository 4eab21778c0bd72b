"""CPU only: how often does a Khosla solve on a FEASIBLE square instance abandon its eps-schedule (sla_stats.restarts)?
Runs oracle/jacobi_model.c -- the bit-exact CPU model of the device rounds (tests assert device == model, restarts
included) -- over the instance families of the tests and of the reference's symmetric bench, and prints one JSON line per
family: instances, feasible ones, restarts among the feasible, rounds with and without the schedule.
    python tests/khosla_restart_stats.py > profiles/r02_khosla_restart_stats.jsonl
(lives under tests/ because it runs the oracle package, which is test infrastructure)"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O          # noqa: E402  (test infrastructure; nothing here is a product path)


def symmetric_instance(n, mean_degree, seed, planted, lo, hi):
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


def family(name, gen, count):
    inst = feas = restarts_feasible = restarts_all = 0
    rounds_s = rounds_p = 0
    for seed in range(count):
        n, rp, c, v = gen(seed)
        s = O.jacobi_model("khosla", n, n, rp, c, v)
        inst += 1
        restarts_all += s["stats"]["restarts"]
        if s["stats"]["num_unassigned"] == 0:
            feas += 1
            restarts_feasible += s["stats"]["restarts"]
            p = O.jacobi_model("khosla", n, n, rp, c, v, khosla_scaling=False)
            assert p["stats"]["num_unassigned"] == 0
            rounds_s += s["stats"]["rounds"]
            rounds_p += p["stats"]["rounds"]
    rec = dict(family=name, instances=inst, feasible=feas, restarts_among_feasible=restarts_feasible,
               restarts_among_all=restarts_all, rounds_feasible_with_schedule=int(rounds_s),
               rounds_feasible_plain=int(rounds_p))
    print(json.dumps(rec), flush=True)
    return rec


def main():
    O.build()

    def bench_shape(seed):            # benches/benchmark.rs:16-47: Bernoulli density, planted permutation, U(500, 1000)
        n = (500, 1000, 2000)[seed % 3]
        return (n,) + symmetric_instance(n, 12, 9000 + seed, True, 500.0, 1000.0)

    def boundary_planted(seed):       # the planted quarter of test_square_khosla_feasibility_boundary_equals_reference
        rng = np.random.default_rng(seed)
        n = int(rng.integers(200, 2001))
        integer = seed % 2 == 0
        rp, c, v = symmetric_instance(n, float(rng.uniform(2.0, 4.0)), 5000 + seed, True, 0.0, 50.0 if integer else 10.0)
        return n, rp, c, np.floor(v) if integer else v

    def boundary_unplanted(seed):     # the unplanted part: a quarter of them happens to have a perfect matching
        rng = np.random.default_rng(seed)
        n = int(rng.integers(10, 201))
        integer = seed % 2 == 0
        rp, c, v = symmetric_instance(n, float(rng.uniform(2.0, 4.0)), 5000 + seed, False, 0.0, 50.0 if integer else 10.0)
        return n, rp, c, np.floor(v) if integer else v

    def cfg4_shape(seed):             # 512 x 512, k = 32 planted, integer costs in [300, 1000)
        from helpers import random_sparse_instance
        rng = np.random.default_rng(700 + seed)
        rp, c, v = random_sparse_instance(rng, 512, 512, 32, integer=True, lo=300, hi=1000)
        return 512, rp, c, v

    family("reference symmetric bench shape (mean degree 12, planted, U(500,1000)), n = 500 / 1000 / 2000", bench_shape, 60)
    family("feasibility boundary, planted (mean degree 2-4, n = 200 .. 2000)", boundary_planted, 80)
    family("feasibility boundary, unplanted (mean degree 2-4, n = 10 .. 200)", boundary_unplanted, 400)
    family("cfg4 shape (512 x 512, k = 32, planted, integer costs)", cfg4_shape, 100)


if __name__ == "__main__":
    main()
