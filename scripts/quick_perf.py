"""Development probe: per-config solve times and the per-round profile (not the bench; see bench.py)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G


def run(name, cls, n, m, k, planted, eps=None, reps=5, options=None):
    solver, z = cls.new(n, m, n * k)
    G.kregular_device(solver, n, m, k, seed=1, planted=planted)
    for k_, v_ in (options or {}).items():
        solver.set_option(k_, v_)
    times = []
    for _ in range(reps):
        t = time.perf_counter()
        st = solver.solve_resident(False, eps)
        times.append((time.perf_counter() - t) * 1e3)
    print(name, options or {}, json.dumps({k_: st[k_] for k_ in ("num_unassigned", "nits", "nreductions", "rounds", "wide_rounds",
          "tail_rounds", "bids", "bid_arcs", "kernel_launches", "graph_launches", "ms_solve")}), "wall_ms", [round(x, 3) for x in times],
          "Garcs/s", round(st["bid_arcs"] / st["ms_solve"] / 1e6, 3), flush=True)
    return solver


def run_cfg4(kind, count=8192):
    b = S.BatchSolver(kind)
    t = time.perf_counter()
    b.generate_device(count, 0, 512, 512, 32, seed=0, planted=True)
    gen = time.perf_counter() - t
    for _ in range(3):
        res = b.solve(eps=None, download=False, per_instance=False)
    tot = res["total"]
    print("cfg4", kind, count, "instances gen_s", round(gen, 3), {k_: tot[k_] for k_ in ("num_unassigned", "rounds", "bids", "bid_arcs",
          "ms_solve")}, "Garcs/s", round(tot["bid_arcs"] / tot["ms_solve"] / 1e6, 3), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg1", "cfg3", "cfg2"]
    if "cfg1" in which:
        run("cfg1 khosla", S.KhoslaSolver, 1000, 10000, 32, False)
    if "cfg3" in which:
        s = run("cfg3 khosla", S.KhoslaSolver, 1_000_000, 4_000_000, 16, False)
        run("cfg3 khosla", S.KhoslaSolver, 1_000_000, 4_000_000, 16, False, options=dict(zero_price_skip=0))
        run("cfg3 khosla", S.KhoslaSolver, 1_000_000, 4_000_000, 16, False, options=dict(graph=0))
        run("cfg3 khosla", S.KhoslaSolver, 1_000_000, 4_000_000, 16, False, options=dict(regular=0))
        run("cfg3 khosla", S.KhoslaSolver, 1_000_000, 4_000_000, 16, False, options=dict(tail_max=512))
        for skip in (1, 0):
            s.set_option("profile", 1)
            s.set_option("zero_price_skip", skip)
            s.solve_resident(False, None)
            for p in s.round_profile():
                b = 12 * p["arcs"] + 8 * p["bidders"]
                print("  skip", skip, p, "bid GB/s", round(b / max(p["bid_ms"], 1e-6) / 1e6, 1) if p["engine"] == 0 else "-")
    if "cfg4" in which:
        run_cfg4("forward")
        run_cfg4("khosla")
    if "cfg2" in which:
        run("cfg2 forward", S.ForwardAuctionSolver, 20000, 20000, 64, True, reps=3)
        run("cfg2 forward", S.ForwardAuctionSolver, 20000, 20000, 64, True, reps=2, options=dict(tail_max=256))
        run("cfg2 forward", S.ForwardAuctionSolver, 20000, 20000, 64, True, reps=2, options=dict(tail_max=512))
        run("cfg2 khosla", S.KhoslaSolver, 20000, 20000, 64, True, reps=2)
