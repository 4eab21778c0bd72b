"""Development probe: round-1 scan with the conflict-resolution atomics removed / confined to an L2-hot range
(variants built with -DSLA_EXP_RED=1/2): how much of the scan do the RED.MAX.64 sector fills cost?  Forward solver with
max_iterations=1 so that the (meaningless) round ends the solve."""
import subprocess, sys, json, os
if len(sys.argv) > 2:
    for so in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, so])
    sys.exit(0)
sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib
so = sys.argv[1]
_lib.LIB_PATH = os.path.abspath(so)
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
n, m, k = 1_000_000, 4_000_000, 16
s, z = S.ForwardAuctionSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
s.set_option("stream_scan", 0)
s.set_option("profile", 1)
bid = []
for _ in range(12):
    s.solve_resident(False, 0.5, max_iterations=1)
    p = s.round_profile()[0]
    bid.append(round(p["bid_ms"] * 1e3, 1))
print(os.path.basename(so), "r1 bid us", bid, flush=True)
