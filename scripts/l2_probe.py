"""Development probe: cfg3 solve and round-1 profile with and without the persisting L2 window on the bid words."""
import sys
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
n, m, k = 1_000_000, 4_000_000, 16
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for _ in range(10):
    s.solve_resident(False, None)
for persist in (0, 1, 0, 1):
    s.set_option("l2_persist", persist)
    for _ in range(5):
        s.solve_resident(False, None)
    ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(15))
    out = {"l2_persist": persist, "ms_solve_median": round(ms[7], 4), "min": round(ms[0], 4)}
    s.set_option("profile", 1)
    bid, asg = [], []
    for _ in range(9):
        s.solve_resident(False, None)
        p = s.round_profile()[0]
        bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
    out["r1_bid_us"] = round(sorted(bid)[4] * 1e3, 1)
    out["r1_assign_us"] = round(sorted(asg)[4] * 1e3, 1)
    s.set_option("profile", 0)
    print(out, flush=True)
