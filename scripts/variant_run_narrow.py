"""Development probe: like variant_run.py, but the instance is uploaded from the host so that the u16 mirror of the
values exists and the uniform-degree scans read it (variants/*.so built by scripts/variant_build.py)."""
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
rp, c, v = G.kregular_host(n, m, k, seed=1)
s.load_csr(n, m, rp, c, v)
s._sync_device()
for _ in range(20):
    st = s.solve_resident(False, None)
ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(9))
out = {"so": os.path.basename(so), "value_bytes": s.scan_value_bytes(), "ms_solve_median": round(ms[4], 4)}
s.set_option("profile", 1)
bid, asg = [], []
for _ in range(9):
    s.solve_resident(False, None)
    p = s.round_profile()[0]
    bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
out["r1_bid_us"] = round(sorted(bid)[4] * 1e3, 1)
out["r1_assign_us"] = round(sorted(asg)[4] * 1e3, 1)
s.set_option("profile_repeat", 8)
bid = []
for _ in range(7):
    s.solve_resident(False, None)
    bid.append(s.round_profile()[0]["bid_ms"] / 8)
out["r1_bid_us_b2b"] = round(sorted(bid)[3] * 1e3, 1)
print(json.dumps(out), flush=True)
