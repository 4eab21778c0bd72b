"""Development probe: builds the library with -DSLA_TAIL_TIMING into a scratch .so and prints where the tail
engine's cycles go on cfg2 (thread 0's view)."""
import ctypes as C
import subprocess
import sys

sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib

so = "/tmp/libsla_timing.so"
extra = sys.argv[1:]
print("flags", extra)
subprocess.run(["nvcc"] + _lib.NVCC_FLAGS + ["-DSLA_TAIL_TIMING"] + extra + ["-o", so, _lib.CSRC + "/sla_api.cu"], check=True)
_lib.LIB_PATH = so
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

import itertools
for (cls, name), sp in itertools.product(((S.ForwardAuctionSolver, "forward"), (S.KhoslaSolver, "khosla")), (1, 0, 2)):
    n = 20000
    solver, z = cls.new(n, n, n * 64)
    G.kregular_device(solver, n, n, 64, seed=1, planted=True)
    solver.set_option("smem_prices", 1 if sp else 0)
    if sp == 2:
        solver.set_option("regular", 0)
    name += f" smem_prices={sp}" + (" regular=0 (no evict_first stream)" if sp == 2 else "")
    st = solver.solve_resident(False, None)
    lib = _lib.load()
    lib.sla_debug_counters.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
    out = (C.c_uint64 * 8)()
    lib.sla_debug_counters(solver._context(), out)
    tot, scan, _, bar1, asg, bar2, rounds, bid = [out[i] for i in range(8)]
    print(name, "ms", st["ms_solve"], "tail rounds", rounds, "cycles/round", tot / max(rounds, 1),
          {k: round(v / max(rounds, 1), 1) for k, v in dict(scan_and_reduce=scan, bid=bid, bar1=bar1, assign=asg, bar2=bar2).items()})
