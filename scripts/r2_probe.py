"""Development probe (round 2): cfg3 / cfg5 resident solve time under option toggles (prune_gather, learn_shape,
prezero_best, narrow mirror of a device-generated CSR), one JSON line per case."""
import json, sys
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G


def run(n, m, k, opts, reps=9, label=""):
    s, z = S.KhoslaSolver.new(n, m, n * k)
    for key, val in opts.items():
        if key in ("narrow_scan",):
            s.set_option(key, val)
    G.kregular_device(s, n, m, k, seed=1)
    for key, val in opts.items():
        s.set_option(key, val)
    for _ in range(4):
        st = s.solve_resident(False, None)
    ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(reps))
    out = dict(label=label, n=n, opts=opts, ms_solve_median=round(ms[len(ms) // 2], 4), ms_min=round(ms[0], 4),
               launches=st["kernel_launches"], rounds=st["rounds"], wide=st["wide_rounds"], vb=s.scan_value_bytes(),
               objective=s.device_objective())
    s.set_option("profile", 1)
    prof = []
    for _ in range(3):
        s.solve_resident(False, None)
        prof = s.round_profile()
    out["profile_us"] = [(p["engine"], p["bidders"], round(p["bid_ms"] * 1e3, 1), round(p["assign_ms"] * 1e3, 1)) for p in prof[:8]]
    print(json.dumps(out), flush=True)
    s.close()


cfg3 = (1_000_000, 4_000_000, 16)
cfg5 = (16_000_000, 64_000_000, 16)
run(*cfg3, {}, label="cfg3 default")
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    run(*cfg5, {}, reps=5, label="cfg5 default")
    sys.exit(0)
run(*cfg3, {"learn_shape": 0}, label="cfg3 no learned shape")
run(*cfg3, {"prune_gather": 0}, label="cfg3 no prune")
run(*cfg3, {"prezero_best": 1}, label="cfg3 prezero")
run(*cfg3, {"narrow_scan": 0}, label="cfg3 f64 scan")
run(*cfg3, {"narrow_scan": 0, "prune_gather": 0}, label="cfg3 f64 scan no prune")
run(*cfg5, {}, reps=5, label="cfg5 default")
run(*cfg5, {"prune_gather": 0}, reps=5, label="cfg5 no prune")
run(*cfg5, {"narrow_scan": 0}, reps=5, label="cfg5 f64 scan")
run(*cfg5, {"narrow_scan": 0, "prune_gather": 0, "learn_shape": 0}, reps=5, label="cfg5 round-1 state")
