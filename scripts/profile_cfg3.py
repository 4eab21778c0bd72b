"""ncu target: cfg3 (Khosla 1M x 4M, k=16) generated in HBM; 2 warm-up solves + 1 solve, host-driven loop so every
round is a separate launch.  argv[1] = zero_price_skip (1/0), argv[2] = stream_scan (1/0)."""
import sys

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

skip = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, m, k = 1_000_000, 4_000_000, 16
solver, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(solver, n, m, k, seed=1)
solver.set_option("graph", 0)
solver.set_option("zero_price_skip", skip)
if len(sys.argv) > 2:
    solver.set_option("stream_scan", int(sys.argv[2]))
for kv in sys.argv[3:]:                      # further options as key=value
    key, val = kv.split("=")
    solver.set_option(key, int(val))
for _ in range(3):
    st = solver.solve_resident(False, None)
print("ok", st["rounds"], st["bid_arcs"], st["ms_solve"])
