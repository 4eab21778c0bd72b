"""Development probe (round 2): where the cycles of a cluster-engine round go.  Takes a library built with -DSLA_MID_TIMING
(scripts/variant_build.py mid_timing "-DSLA_MID_TIMING") and prints, for cfg3 with cluster_engine = 1, the
per-round averages of thread 0 / cluster rank 0: bid phase, barrier 1, resolve, local bookkeeping, barrier 2."""
import ctypes as C, json, subprocess, sys
sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib
import os
so = sys.argv[1] if len(sys.argv) > 1 else "variants/mid_timing.so"     # built here: scripts/variant_build.py mid_timing "-DSLA_MID_TIMING"
extra = [os.path.basename(so)]
_lib.LIB_PATH = os.path.abspath(so)
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
lib = _lib.load()
lib.sla_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
for label, (n, m, k) in (("cfg3", (1_000_000, 4_000_000, 16)),):
    for handover in (24, 1):
        s, z = S.KhoslaSolver.new(n, m, n * k)
        G.kregular_device(s, n, m, k, seed=1)
        s.set_option("cluster_engine", 1)
        s.set_option("cluster_handover", handover)
        for _ in range(6):
            st = s.solve_resident(False, None)
        ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(9))
        out = (C.c_uint64 * 24)()
        lib.sla_debug_counters(s._context(), out)
        r = max(out[6], 1)
        seg = ["bid (row, prices, bid, atomicMax issue)", "cluster barrier 1", "resolve (word, owner, stores issue)", "block sync + lengths", "cluster barrier 2"]
        print(json.dumps({"cfg": label, "flags": extra, "handover": handover, "ms_solve_median": round(ms[4], 4), "cluster_rounds": int(out[6]),
                          "cycles_per_round": {seg[i]: round(out[i] / r, 1) for i in range(5)},
                          "sum_cycles_per_round": round(sum(out[0:5]) / r, 1),
                          "first_rounds_bidders_cycles": [(int(out[9 + 2 * i]), int(out[8 + 2 * i])) for i in range(8)], "stats": {k_: st[k_] for k_ in ("rounds", "wide_rounds", "tail_rounds", "cluster_rounds")}}), flush=True)
        s.close()
