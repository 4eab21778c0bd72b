"""Development probe: one long eviction chain through the tail engine's small rounds (3 persons fight over 2 good
objects in steps of eps: ~1e5 rounds with a single bidder), for ncu's per-instruction stall sampling."""
import sys

import numpy as np

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S

eps = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
solver, z = S.KhoslaSolver.new(3, 3, 9)
solver.init(3, 3)
for i in range(3):
    solver.extend_from_values(i, np.array([0, 1, 2], dtype=np.uint32), np.array([1000.0, 1000.0, 0.0]))
solver.solve(z, True, eps)
st = solver.last_stats
print("chain", st["tail_rounds"], "rounds", round(st["ms_solve"], 3), "ms", round(st["ms_solve"] * 1.965e6 / st["tail_rounds"], 1),
      "cycles/round", "unassigned", z.num_unassigned)
