"""Development probe: per-round profile and solve time of prebuilt library variants at cfg3."""
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
    s.solve_resident(False, None)
ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(15))
s.set_option("profile", 1)
acc = {}
for _ in range(7):
    s.solve_resident(False, None)
    for p in s.round_profile():
        acc.setdefault(p["round"], []).append((p["bid_ms"], p["assign_ms"]))
print(os.path.basename(so), "ms_solve median", round(ms[7], 4), "rounds (bid us, assign us):",
      {r: (round(sorted(x[0] for x in v)[3] * 1e3, 1), round(sorted(x[1] for x in v)[3] * 1e3, 1)) for r, v in acc.items()}, flush=True)
