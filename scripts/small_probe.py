"""Development probe: Khosla on a planted symmetric k=64 instance small enough that the whole solve is one launch of
the tail engine with ~2 bidders per round (for ncu's per-instruction stall sampling of the small rounds)."""
import sys

sys.path.insert(0, ".")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
cls = S.ForwardAuctionSolver if (len(sys.argv) > 2 and sys.argv[2] == "forward") else S.KhoslaSolver
solver, z = cls.new(n, n, n * 64)
G.kregular_device(solver, n, n, 64, seed=1, planted=True)
st = solver.solve_resident(False, None)
print(cls.__name__, n, {k: st[k] for k in ("rounds", "tail_rounds", "bids", "kernel_launches", "graph_launches", "num_unassigned")},
      round(st["ms_solve"], 3), "ms", round(st["ms_solve"] * 1.965e6 / max(st["tail_rounds"], 1), 1), "cycles/tail round")
