"""Development probe: host-side passes on the GPU box's cores (in-place negation of 16M f64 by the library's helper,
numpy narrowing casts): how fast can the host prepare a cfg3 upload?"""
import ctypes as C, sys, time
sys.path.insert(0, ".")
import numpy as np
from sparse_linear_assignment_b200 import _lib, solver as SV
lib = _lib.load()
n = 16_000_000
v = SV.host_array(n, np.float64)
v[:] = np.random.default_rng(0).integers(300, 1000, size=n).astype(np.float64)
for th in (1, 2, 4, 8, 16):
    ts = []
    for _ in range(7):
        t = time.perf_counter(); lib.sla_host_negate_f64(v.ctypes.data, n, th); ts.append(time.perf_counter() - t)
    print("negate in place, threads", th, "median ms", round(sorted(ts)[3] * 1e3, 3), "GB/s (r+w)", round(2 * 8 * n / sorted(ts)[3] / 1e9, 1), flush=True)
u = np.empty(n, dtype=np.uint16)
ts = []
for _ in range(5):
    t = time.perf_counter(); np.copyto(u, v, casting="unsafe"); ts.append(time.perf_counter() - t)
print("numpy f64->u16 single thread ms", round(sorted(ts)[2] * 1e3, 2))
