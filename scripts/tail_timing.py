"""Development probe: builds the library with -DSLA_TAIL_TIMING into a scratch .so and prints where the cycles of the
tail engine's small rounds go (stamps of the busiest warp; see TK() in csrc/sla_kernels.cuh)."""
import ctypes as C
import subprocess
import sys

import numpy as np

sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib

so = "/tmp/libsla_timing.so"
extra = [a for a in sys.argv[1:] if a.startswith("-") and a != "--notiming"]
TIMING = "--notiming" not in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["chain", "cfg2f", "cfg2k"]
print("flags", extra)
subprocess.run(["nvcc"] + _lib.NVCC_FLAGS + (["-DSLA_TAIL_TIMING"] if TIMING else []) + extra + ["-o", so, _lib.CSRC + "/sla_api.cu"], check=True)
_lib.LIB_PATH = so
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

SEG = ["lane scan (price + keys)", "REDUX agreement", "owner", "prefetch issue + bid", "STS", "barrier 1", "resolve (LDS + any)",
       "stores + row move", "still-active flag", "barrier 2", "flag vote + loop"]


def report(name, solver, st):
    lib = _lib.load()
    lib.sla_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    out = (C.c_uint64 * 24)()
    lib.sla_debug_counters(solver._context(), out)
    if not TIMING:
        print(name, "ms", round(st["ms_solve"], 3), "tail rounds", st["tail_rounds"], "cycles/round (1.965 GHz)",
              round(st["ms_solve"] * 1.965e6 / max(st["tail_rounds"], 1), 1))
        return
    tot, rounds, act = out[0], max(out[6], 1), max(out[7], 1)
    seg = {SEG[i]: round(out[8 + i] / act, 1) for i in range(11)}
    print(name, "ms", round(st["ms_solve"], 3), "tail rounds", out[6], "cycles/round", round(tot / rounds, 1),
          "busiest warp active rounds", out[7], "its cycles/active round", round(sum(out[8:19]) / act, 1))
    for k, v in seg.items():
        print(f"      {k:28s} {v}")


if "chain" in which:
    # one long eviction chain: 3 persons fight over 2 good objects in steps of eps -> ~2e5 rounds with one bidder
    for cls, name in ((S.KhoslaSolver, "chain khosla 3x3"),):
        solver, z = cls.new(3, 3, 9)
        solver.init(3, 3)
        for i in range(3):
            solver.extend_from_values(i, np.array([0, 1, 2], dtype=np.uint32), np.array([1000.0, 1000.0, 0.0]))
        for own in (1, 0):
            solver.set_option("smem_owners", own)
            solver._dirty = True
            solver.solve(z, True, 0.01)
            report(f"{name} smem_owners={own}", solver, solver.last_stats)

for tag, cls in (("cfg2f", S.ForwardAuctionSolver), ("cfg2k", S.KhoslaSolver)):
    if tag not in which:
        continue
    for sp in (3, 1, 0):
        n = 20000
        solver, z = cls.new(n, n, n * 64)
        G.kregular_device(solver, n, n, 64, seed=1, planted=True)
        solver.set_option("smem_prices", 1 if sp else 0)
        solver.set_option("smem_owners", 1 if sp == 3 else 0)
        st = solver.solve_resident(False, None)
        report(f"{tag} smem_prices={1 if sp else 0} smem_owners={1 if sp == 3 else 0}", solver, st)

if "cfg3" in which:
    n, m, k = 1_000_000, 4_000_000, 16
    solver, z = S.KhoslaSolver.new(n, m, n * k)
    G.kregular_device(solver, n, m, k, seed=1)
    for tm in (1024, 256, 64, 24):
        solver.set_option("tail_max", tm)
        for _ in range(3):
            st = solver.solve_resident(False, None)
        report(f"cfg3 tail_max={tm} (solve {st['ms_solve']:.4f} ms, wide {st['wide_rounds']}, tail {st['tail_rounds']})", solver, st)
