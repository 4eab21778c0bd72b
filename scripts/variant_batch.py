"""Development probe: cfg4 batch engine built with extra nvcc flags."""
import subprocess
import sys

sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib

extra = sys.argv[1:]
so = "/tmp/libsla_variant.so"
subprocess.run(["nvcc"] + _lib.NVCC_FLAGS + extra + ["-o", so, _lib.CSRC + "/sla_api.cu"], check=True)
_lib.LIB_PATH = so
import sparse_linear_assignment_b200 as S

for kind in ("forward", "khosla"):
    b = S.BatchSolver(kind)
    b.generate_device(8192, 0, 512, 512, 32, seed=0, planted=True)
    for _ in range(3):
        tot = b.solve(download=False, per_instance=False)["total"]
    print(extra, kind, "ms", round(tot["ms_solve"], 2), "rounds", tot["rounds"], "unassigned", tot["num_unassigned"], flush=True)
