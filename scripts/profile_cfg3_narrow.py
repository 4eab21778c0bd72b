"""ncu target: cfg3 (Khosla 1M x 4M, k=16) uploaded from the host, so that the u16 copy of the values stays in HBM and
the uniform-degree scans read it (sla_scan_value_bytes == 2); host-driven loop, every round a separate launch.
argv[1] = zero_price_skip (1/0), further options as key=value."""
import sys

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

skip = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, m, k = 1_000_000, 4_000_000, 16
solver, z = S.KhoslaSolver.new(n, m, n * k)
rp, c, v = G.kregular_host(n, m, k, seed=1)
solver.load_csr(n, m, rp, c, v)
solver.set_option("graph", 0)
solver.set_option("zero_price_skip", skip)
for kv in sys.argv[2:]:
    key, val = kv.split("=")
    solver.set_option(key, int(val))
solver._sync_device()
for _ in range(3):
    st = solver.solve_resident(False, None)
print("ok", solver.scan_value_bytes(), st["rounds"], st["bid_arcs"], st["ms_solve"])
