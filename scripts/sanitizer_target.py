"""compute-sanitizer target: small solves through every engine (wide + tail, Forward phases, batch, partitioned)."""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G
from sparse_linear_assignment_b200.distributed import CudaShardEngine, PartitionedKhoslaSolver
from helpers import random_sparse_instance

rng = np.random.default_rng(0)
# wide + tail, regular CSR (k % 8 == 0) and ragged CSR
for cls, n, m, k in ((S.KhoslaSolver, 3000, 5000, 16), (S.ForwardAuctionSolver, 1500, 1500, 24), (S.KhoslaSolver, 2000, 2500, 7)):
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=0, hi=300)
    s, z = cls.new(n, m, n * k)
    s.load_csr(n, m, rp, c, v)
    s.solve(z, False, 1.0 / (m + 1))
    print(cls.__name__, n, m, k, "objective", s.get_objective(z), "unassigned", z.num_unassigned, s.device_validate_matching())
# batch
b = S.BatchSolver("forward")
b.generate_device(16, 0, 128, 128, 16, seed=0, planted=True)
print("batch", b.solve()["total"]["num_unassigned"])
# partitioned, both exchanges, world 1
for ex in ("dense", "sparse"):
    rp, c, v = random_sparse_instance(rng, 1200, 2000, 8, integer=True, lo=1, hi=100)
    s, _ = S.KhoslaSolver.new(1200, 2000, 1200 * 8)
    s.load_csr(1200, 2000, rp, c, v)
    r = PartitionedKhoslaSolver(CudaShardEngine(s), exchange=ex).solve(False, None)
    print("partitioned", ex, r["stats"]["global_num_unassigned"], r["stats"]["rounds"])
print("sanitizer target done")
