"""Development probe: event-timed duration of R back-to-back launches of the round-1 scan (profile_repeat): the slope is
the kernel's own duration, the intercept the overhead of bracketing ONE launch with CUDA events."""
import sys
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
n, m, k = 1_000_000, 4_000_000, 16
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for ss in (0, 1):
    s.set_option("stream_scan", ss)
    s.set_option("profile", 1)
    for rep in (1, 2, 4, 8, 1):
        s.set_option("profile_repeat", rep)
        t = []
        for _ in range(9):
            s.solve_resident(False, None)
            t.append(s.round_profile()[0]["bid_ms"] * 1e3)
        t.sort()
        print("stream_scan", ss, "repeat", rep, "median us", round(t[4], 1), "per launch", round(t[4] / rep, 1), flush=True)
    s.set_option("profile_repeat", 1)
    s.set_option("profile", 0)
