"""Development probe: build the library with extra nvcc flags into a scratch .so and print the cfg3 round profile."""
import subprocess
import sys

sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib

extra = sys.argv[1:]
so = "/tmp/libsla_variant.so"
subprocess.run(["nvcc"] + _lib.NVCC_FLAGS + extra + ["-o", so, _lib.CSRC + "/sla_api.cu"], check=True)
_lib.LIB_PATH = so
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

n, m, k = 1_000_000, 4_000_000, 16
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for _ in range(3):
    st = s.solve_resident(False, None)
ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(7))
out = {"flags": extra, "ms_solve_median": round(ms[3], 4)}
for skip in (1, 0):
    s.set_option("profile", 1)
    s.set_option("zero_price_skip", skip)
    bid, asg = [], []
    for _ in range(5):
        s.solve_resident(False, None)
        p = s.round_profile()[0]
        bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
    out[f"r1_bid_us_skip{skip}"] = round(sorted(bid)[2] * 1e3, 1)
    out[f"r1_assign_us_skip{skip}"] = round(sorted(asg)[2] * 1e3, 1)
    s.set_option("profile", 0)
print(out)
