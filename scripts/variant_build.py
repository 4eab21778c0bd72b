"""Development probe: build library variants in parallel: variant_build.py name1 "flags1" name2 "flags2" ..."""
import subprocess, sys, os
sys.path.insert(0, ".")
from sparse_linear_assignment_b200 import _lib
os.makedirs("variants", exist_ok=True)
procs = []
args = sys.argv[1:]
for name, flags in zip(args[0::2], args[1::2]):
    so = f"variants/{name}.so"
    procs.append((name, subprocess.Popen(["nvcc"] + _lib.NVCC_FLAGS + flags.split() + ["-o", so, _lib.CSRC + "/sla_api.cu"])))
for name, p in procs:
    print(name, "rc", p.wait())
