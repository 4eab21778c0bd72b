"""GPU probe (not a pytest file; lives under tests/ because the oracle is its checker): cfg3's shape end to end through
KhoslaSolver.solve() with host buffers for the three wire formats of `values` --
    integers (u16 on the wire, u16 mirror read by the scans: what bench.py times),
    integer + 0.5 (exact in f32: 4 bytes on the wire, scans read f64),
    integer + uniform fraction (real-valued weights: f64 on the wire, scans read f64)
-- with the device time of the solve, the bytes the upload moved, and the reference bar checked against the CPU port
(same num_unassigned; objective bit-exact for the integer class, within n * eps for the others).
    python tests/probe_real_valued_e2e.py > gpurun_out/real_valued_e2e.jsonl"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sparse_linear_assignment_b200 as S                       # noqa: E402
from sparse_linear_assignment_b200 import generators as G       # noqa: E402
from oracle import oracle as O                                   # noqa: E402


def main():
    n, m, k = 1_000_000, 4_000_000, 16
    rp, c, v = G.kregular_host(n, m, k, seed=1)
    rng = np.random.default_rng(7)
    classes = (("integers in [300, 1000)", v, True),
               ("integer + 0.5", v + 0.5, False),
               ("integer + uniform fraction", v + rng.random(v.size), False))
    for name, vals, integer in classes:
        solver, z = S.KhoslaSolver.new(n, m, n * k)
        solver.load_csr(n, m, rp, c, vals)
        hv = solver.values()
        e2e, dev = [], []
        for it in range(3 + 9):
            if hv[0] < 0:
                np.negative(hv, out=hv)          # undo the in-place sign normalisation of the previous solve
            solver._dirty = True                 # the host CSR "changed": upload again, as a fresh problem would
            t = time.perf_counter()
            solver.solve(z, False, None)
            dt = time.perf_counter() - t
            if it >= 3:
                e2e.append(dt * 1e3)
                dev.append(float(solver.last_stats["ms_solve"]))
        moved, width = solver.last_upload()
        obj = solver.get_objective(z)
        o = O.OracleSolver("khosla", n, m, len(c))
        o.load_csr(n, m, rp, c, vals)
        t = time.perf_counter()
        o.solve()
        cpu_ms = (time.perf_counter() - t) * 1e3
        ref = o.get_objective()
        ok = (z.num_unassigned == o.num_unassigned) and (obj == ref if integer else abs(obj - ref) <= n * z.eps)
        e2e.sort()
        dev.sort()
        print(json.dumps({"values": name, "wire_bytes_per_value": width, "h2d_bytes": moved,
                          "scan_value_bytes": solver.scan_value_bytes(), "e2e_ms_median": round(e2e[len(e2e) // 2], 3),
                          "e2e_ms_min": round(e2e[0], 3), "ms_solve_median": round(dev[len(dev) // 2], 4),
                          "rounds": int(solver.last_stats["rounds"]), "bid_arcs": int(solver.last_stats["bid_arcs"]),
                          "objective": obj, "cpu_port_objective": ref, "cpu_port_ms": round(cpu_ms, 1),
                          "num_unassigned": int(z.num_unassigned), "reference_bar_met": bool(ok)}), flush=True)
        solver.close()
        del solver, z, hv, o


if __name__ == "__main__":
    main()
