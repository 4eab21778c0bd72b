"""Development probe: first-round scan through the TMA pipeline (stream_scan=1) against the LDG.256 kernel (0):
round-1 profile, whole-solve time, and equality of the results."""
import sys, json
sys.path.insert(0, ".")
import numpy as np
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

def run(n, m, k, cls=S.KhoslaSolver, planted=False, eps=None):
    s, z = cls.new(n, m, n * k)
    G.kregular_device(s, n, m, k, seed=1, planted=planted)
    res = {}
    for mode in (0, 1, 0, 1):
        s.set_option("stream_scan", mode)
        for _ in range(10):
            s.solve_resident(False, eps)
        ms = sorted(s.solve_resident(False, eps)["ms_solve"] for _ in range(9))
        s.set_option("profile", 1)
        bid, asg = [], []
        for _ in range(7):
            s.solve_resident(False, eps)
            p = s.round_profile()[0]
            bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
        s.set_option("profile", 0)
        st = s.solve_resident(False, eps)
        s.download_solution(z)
        res[mode] = (z.person_to_object.copy(), s.prices().copy(), st["rounds"], st["bids"])
        alg = 12 * n * k + 8 * n
        b = sorted(bid)[3]
        print(json.dumps({"case": f"{cls.__name__} {n}x{m} k={k}", "stream_scan": mode, "ms_solve_median": round(ms[4], 4),
                          "r1_bid_us": round(b * 1e3, 1), "GB/s": round(alg / b / 1e6, 0), "r1_assign_us": round(sorted(asg)[3] * 1e3, 1),
                          "rounds": st["rounds"], "unassigned": st["num_unassigned"]}), flush=True)
    same = np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1]) and res[0][2:] == res[1][2:]
    print("  results identical between the two scans:", same, flush=True)
    assert same

run(1_000_000, 4_000_000, 16)
run(1_000_003, 4_000_000, 24)
run(300_000, 1_000_000, 64)
run(100_001, 400_000, 8)
run(50_000, 50_000, 256, cls=S.ForwardAuctionSolver, planted=True, eps=None)
if len(sys.argv) > 1:
    run(16_000_000, 64_000_000, 16)
