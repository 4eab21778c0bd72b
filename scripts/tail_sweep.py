"""Development probe: cfg3 solve time against the bidder count at which the single-CTA tail engine takes over."""
import sys
sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
n, m, k = 1_000_000, 4_000_000, 16
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
for _ in range(10):
    s.solve_resident(False, None)
import itertools
sr_list = [int(x) for x in sys.argv[1:]] or [6]
for sr, tm in itertools.product(sr_list, (1024, 512, 256, 128, 64, 32, 16, 8)):
    s.set_option("super_rounds", sr)
    s.set_option("tail_max", tm)
    for _ in range(3):
        s.solve_resident(False, None)
    r = sorted((s.solve_resident(False, None) for _ in range(11)), key=lambda d: d["ms_solve"])[5]
    print("super_rounds", sr, "tail_max", tm, "ms_solve", round(r["ms_solve"], 4), "wide", r["wide_rounds"], "tail", r["tail_rounds"], "launches", r["kernel_launches"], "graphs", r["graph_launches"], flush=True)
s.set_option("tail_max", 1024)
s.set_option("super_rounds", 6)
s.set_option("profile", 1)
s.solve_resident(False, None)
for p in s.round_profile():
    print(p)
