"""ncu target: one cfg2 Forward solve (tail engine) or one cfg4 batch solve of 1024 instances (batch engine)."""
import sys

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

which = sys.argv[1]
if which == "tail":
    n = 20000
    s, z = S.ForwardAuctionSolver.new(n, n, n * 64)
    G.kregular_device(s, n, n, 64, seed=1, planted=True)
    s.set_option("graph", 0)
    st = s.solve_resident(False, None)
    print("ok", st["rounds"], st["tail_rounds"], st["ms_solve"])
else:
    b = S.BatchSolver("forward")
    b.generate_device(2048, 0, 512, 512, 32, seed=0, planted=True)
    tot = b.solve(download=False, per_instance=False)["total"]
    print("ok", tot["rounds"], tot["ms_solve"])
