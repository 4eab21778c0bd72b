"""Development probe (round 2): cfg3 end to end (KhoslaSolver.solve with host buffers) with the statistics and the
widening riding behind the u16 pieces (option upload_ride = 1, default) and as two passes at the end (0), interleaved."""
import json, sys, time
sys.path.insert(0, ".")
import numpy as np
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

n, m, k = 1_000_000, 4_000_000, 16
rp, c, v = G.kregular_host(n, m, k, seed=1)
solver, z = S.KhoslaSolver.new(n, m, n * k)
solver.load_csr(n, m, rp, c, v)
hv = solver.values()
res = {0: [], 1: []}
for it in range(4 + 24):
    ride = it & 1
    solver.set_option("upload_ride", ride)
    if hv[0] < 0:
        np.negative(hv, out=hv)
    solver._dirty = True
    t = time.perf_counter()
    solver.solve(z, False, None)
    dt = time.perf_counter() - t
    if it >= 4:
        res[ride].append(dt * 1e3)
obj = solver.get_objective(z)
for ride in (1, 0):
    xs = sorted(res[ride])
    print(json.dumps({"upload_ride": ride, "e2e_ms_median": round(xs[len(xs) // 2], 3), "min": round(xs[0], 3), "max": round(xs[-1], 3),
                      "objective": obj, "scan_value_bytes": solver.scan_value_bytes()}), flush=True)
