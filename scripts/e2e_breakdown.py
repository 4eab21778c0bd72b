"""Development probe: where the end-to-end time of KhoslaSolver.solve() goes at cfg3 (host buffers page-locked)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import _lib, generators as G

n, m, k = 1_000_000, 4_000_000, 16
rp, c, v = G.kregular_host(n, m, k, seed=1)
solver, z = S.KhoslaSolver.new(n, m, n * k)
solver.load_csr(n, m, rp, c, v)
lib = _lib.load()
ctx = solver._context()
hv = solver.values()
rpp = np.ascontiguousarray(solver.i_starts_stops()[: n + 1], dtype=np.uint32)
cc = solver.column_indices()


def t(f, reps=5):
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        f()
        best = min(best, time.perf_counter() - t0)
    return round(best * 1e3, 3)


print("plain upload ms        ", t(lambda: _lib.check(ctx, lib.sla_upload_csr(ctx, n, m, rpp.ctypes.data, cc.ctypes.data, hv.ctypes.data, n * k))))
print("negating upload ms (8) ", t(lambda: _lib.check(ctx, lib.sla_upload_csr_negating(ctx, n, m, rpp.ctypes.data, cc.ctypes.data, hv.ctypes.data, n * k, 8))))
print("negating upload ms (4) ", t(lambda: _lib.check(ctx, lib.sla_upload_csr_negating(ctx, n, m, rpp.ctypes.data, cc.ctypes.data, hv.ctypes.data, n * k, 4))))
print("host negate only ms (8)", t(lambda: lib.sla_host_negate_f64(hv.ctypes.data, hv.size, 8)))
print("host negate only ms (1)", t(lambda: lib.sla_host_negate_f64(hv.ctypes.data, hv.size, 1)))
if hv[0] < 0:
    np.negative(hv, out=hv)
p2o = S.solver.host_array(n, np.uint32)
o2p = S.solver.host_array(m, np.uint32)
st = _lib.SlaStats()
print("solve + D2H(p2o,o2p) ms", t(lambda: _lib.check(ctx, lib.sla_khosla_solve(ctx, 0, float("nan"), p2o.ctypes.data, o2p.ctypes.data, None, C.byref(st)))))
print("solve resident ms      ", t(lambda: _lib.check(ctx, lib.sla_khosla_solve(ctx, 0, float("nan"), None, None, None, C.byref(st)))))


def full():
    if hv[0] < 0:
        np.negative(hv, out=hv)
    solver._dirty = True
    t0 = time.perf_counter()
    solver.solve(z, False, None)
    return time.perf_counter() - t0


print("solver.solve() e2e ms  ", round(min(full() for _ in range(5)) * 1e3, 3))
