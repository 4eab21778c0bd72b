"""Development probe: wall-clock split of KhoslaSolver.solve() at cfg3 (upload / solve call / wrapper)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G, solver as SV

n, m, k = (1_000, 10_000, 32) if (len(sys.argv) > 1 and sys.argv[1] == "cfg1") else (1_000_000, 4_000_000, 16)
rp, c, v = G.kregular_host(n, m, k, seed=1)
solver, z = S.KhoslaSolver.new(n, m, n * k)
solver.load_csr(n, m, rp, c, v)
hv = solver.values()
marks = {}


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **kw):
        t = time.perf_counter()
        r = f(*a, **kw)
        marks[name] = marks.get(name, 0.0) + time.perf_counter() - t
        return r
    setattr(obj, name, g)


for name in ("_sync_device", "_outputs", "_begin_negation", "_finish", "validate_input"):
    wrap(solver, name)
lib = SV._lib.load()
orig = lib.sla_khosla_solve


def timed_solve(*a):
    t = time.perf_counter()
    r = orig(*a)
    marks["sla_khosla_solve"] = marks.get("sla_khosla_solve", 0.0) + time.perf_counter() - t
    return r


class LibProxy:
    def __getattr__(self, k):
        return timed_solve if k == "sla_khosla_solve" else getattr(lib, k)


SV._lib.load = lambda: LibProxy()
reps = 200 if n < 10_000 else 8
tot = 0.0
for it in range(reps + 2):
    if hv[0] < 0:
        np.negative(hv, out=hv)
    solver._dirty = True
    if it == 2:
        marks.clear()
        tot = 0.0
    t = time.perf_counter()
    solver.solve(z, False, None)
    tot += time.perf_counter() - t
print("e2e ms", round(tot / reps * 1e3, 3), {k_: round(v_ / reps * 1e3, 3) for k_, v_ in marks.items()}, "host threads", SV.host_threads())
