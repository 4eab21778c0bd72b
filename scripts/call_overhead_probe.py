"""Development probe: per-solve wall clock through solve_resident (Python mirror) against a bare ctypes loop over
sla_khosla_solve, cfg3 resident: how much of the step is the wrapper?"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G, _lib
n, m, k = 1_000_000, 4_000_000, 16
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for _ in range(20):
    st = s.solve_resident(False, None)
R = 200
t = time.perf_counter()
for _ in range(R):
    st = s.solve_resident(False, None)
w1 = (time.perf_counter() - t) / R * 1e3
lib = _lib.load()
ctx = s._context()
stats = _lib.SlaStats()
nan = float("nan")
f = lib.sla_khosla_solve
ref = C.byref(stats)
t = time.perf_counter()
for _ in range(R):
    f(ctx, 0, nan, None, None, None, ref)
w2 = (time.perf_counter() - t) / R * 1e3
print("solve_resident wall ms", round(w1, 4), "bare ctypes wall ms", round(w2, 4), "ms_solve", round(st["ms_solve"], 4), "->", round(stats.ms_solve, 4))
