import sys, time, json
sys.path.insert(0, ".")
import numpy as np, torch
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def run(n, m, k, opts, cls=S.KhoslaSolver, planted=False, eps=None):
    solver, z = cls.new(n, m, n * k)
    G.kregular_device(solver, n, m, k, seed=1, planted=planted)
    for a, b in opts.items(): solver.set_option(a, b)
    res = []
    for _ in range(8):
        flush.zero_(); torch.cuda.synchronize()
        t = time.perf_counter(); st = solver.solve_resident(False, eps); w = (time.perf_counter() - t) * 1e3
        res.append((st["ms_solve"], w))
    res.sort()
    print(cls.__name__, n, m, k, opts, "ms_solve med %.4f min %.4f wall med %.4f" % (res[4][0], res[0][0], sorted(r[1] for r in res)[4]),
          {k_: st[k_] for k_ in ("rounds", "wide_rounds", "tail_rounds", "kernel_launches", "graph_launches", "bids")}, flush=True)
for n, m, k in ((1000, 10000, 32), (100, 1000, 32), (500, 5000, 32), (1900, 19000, 32)):
    for tm in (1024, 512, 256, 128, 64):
        run(n, m, k, dict(tail_max=tm))
