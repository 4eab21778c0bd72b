"""Development probe (round 2): resident solve time with the cluster engine on / off and at several hand-over points,
cfg3 and cfg5, one JSON line per case."""
import json, sys
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G


def run(n, m, k, opts, reps=9, label=""):
    s, z = S.KhoslaSolver.new(n, m, n * k)
    G.kregular_device(s, n, m, k, seed=1)
    for key, val in opts.items():
        s.set_option(key, val)
    for _ in range(4):
        st = s.solve_resident(False, None)
    ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(reps))
    out = dict(label=label, n=n, opts=opts, ms_solve_median=round(ms[len(ms) // 2], 4), ms_min=round(ms[0], 4),
               launches=st["kernel_launches"], rounds=st["rounds"], wide=st["wide_rounds"], tail=st["tail_rounds"],
               cluster=st["cluster_rounds"], objective=s.device_objective())
    s.set_option("profile", 1)
    prof = []
    for _ in range(3):
        s.solve_resident(False, None)
        prof = s.round_profile()
    out["profile_us"] = [(p["engine"], p["bidders"], p["rounds_covered"], round(p["bid_ms"] * 1e3, 1), round(p["assign_ms"] * 1e3, 1))
                         for p in prof[:12]]
    print(json.dumps(out), flush=True)
    s.close()


cfg3 = (1_000_000, 4_000_000, 16)
cfg5 = (16_000_000, 64_000_000, 16)
run(*cfg3, {"cluster_engine": 1}, label="cfg3 with the cluster engine")
run(*cfg3, {"cluster_engine": 0}, label="cfg3 without the cluster engine")
for h in (1, 64, 256):
    run(*cfg3, {"cluster_engine": 1, "cluster_handover": h}, label=f"cfg3 hand-over at {h}")
run(*cfg5, {"cluster_engine": 1}, reps=5, label="cfg5 with the cluster engine")
run(*cfg5, {"cluster_engine": 0}, reps=5, label="cfg5 without the cluster engine")
