"""Development probe: cfg3 solve and round-1 profile with and without the L2 access-policy window on the object state."""
import ctypes, sys
sys.path.insert(0, ".")
import torch
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G

rt = ctypes.CDLL("libcudart.so.12")
for name, attr in (("maxPersistingL2", 108), ("maxAccessPolicyWindow", 109), ("l2CacheSize", 38)):
    v = ctypes.c_int()
    rt.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0)
    print(name, v.value)

n, m, k = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (1_000_000, 4_000_000, 16)))
s, z = S.KhoslaSolver.new(n, m, n * k)
G.kregular_device(s, n, m, k, seed=1)
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
def pre(mode):
    if mode == "clean":
        big.sum()
    elif mode == "dirty":
        big.zero_()
    torch.cuda.synchronize()
for _ in range(5):
    s.solve_resident(False, None)
for mode in ("none", "clean", "dirty", "none", "clean", "dirty"):
    s.set_option("profile", 1)
    bid, asg = [], []
    for _ in range(5):
        pre(mode)
        s.solve_resident(False, None)
        p = s.round_profile()[0]
        bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
    s.set_option("profile", 0)
    print("pre-solve L2 state", mode, "r1 bid us", [round(b * 1e3, 1) for b in bid], "assign", [round(b * 1e3, 1) for b in asg], flush=True)
for persist in (0, 1, 0, 1):
    s.set_option("l2_persist", persist)
    for _ in range(3):
        s.solve_resident(False, None)
    ms = sorted(s.solve_resident(False, None)["ms_solve"] for _ in range(9))
    out = {"l2_persist": persist, "ms_solve_median": round(ms[4], 4), "min": round(ms[0], 4)}
    for skip in (1, 0):
        s.set_option("profile", 1)
        s.set_option("zero_price_skip", skip)
        bid, asg = [], []
        for _ in range(5):
            s.solve_resident(False, None)
            p = s.round_profile()[0]
            bid.append(p["bid_ms"]); asg.append(p["assign_ms"])
        out[f"r1_bid_us_skip{skip}"] = round(sorted(bid)[2] * 1e3, 1)
        out[f"r1_assign_us_skip{skip}"] = round(sorted(asg)[2] * 1e3, 1)
        s.set_option("profile", 0)
        s.set_option("zero_price_skip", 1)
    print(out, flush=True)
