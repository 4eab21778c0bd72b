"""Development probe: cfg3 solve time and round-1 profile for prebuilt library variants (variants/*.so, built by
scripts/variant_build.py with extra nvcc flags)."""
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
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for _ in range(20):
    st = s.solve_resident(False, None)
ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(9))
out = {"so": os.path.basename(so), "ms_solve_median": round(ms[4], 4)}
for skip in (1, 0):
    s.set_option("profile", 1)
    s.set_option("zero_price_skip", skip)
    bid, asg = [], []
    for _ in range(7):
        s.solve_resident(False, None)
        p = s.round_profile()[0]
        bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
    out[f"r1_bid_us_skip{skip}"] = round(sorted(bid)[3] * 1e3, 1)
    out[f"r1_assign_us_skip{skip}"] = round(sorted(asg)[3] * 1e3, 1)
    s.set_option("profile", 0)
print(json.dumps(out), flush=True)
