"""Development probe (round 2): resident solve time and the per-round event brackets of the first rounds for prebuilt
library variants (variants/*.so from scripts/variant_build.py), cfg3 and cfg5; one JSON line per variant and config."""
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

for label, (n, m, k), warm, reps in (("cfg3", (1_000_000, 4_000_000, 16), 12, 11), ("cfg5", (16_000_000, 64_000_000, 16), 3, 5)):
    s, z = S.KhoslaSolver.new(n, m, n * k)
    G.kregular_device(s, n, m, k, seed=1)
    for _ in range(warm):
        st = s.solve_resident(False, None)
    ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(reps))
    out = {"so": os.path.basename(so), "cfg": label, "ms_solve_median": round(ms[len(ms) // 2], 4), "ms_min": round(ms[0], 4),
           "objective": s.device_objective(), "rounds": st["rounds"], "bids": st["bids"]}
    s.set_option("profile", 1)
    prof = []
    for _ in range(3):
        s.solve_resident(False, None)
        prof = s.round_profile()
    out["profile_us"] = [(p["bidders"], round(p["bid_ms"] * 1e3, 1), round(p["assign_ms"] * 1e3, 1)) for p in prof[:4]]
    print(json.dumps(out), flush=True)
    s.close()
