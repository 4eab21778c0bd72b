"""Development probe (round 2): where the end-to-end time of a small solve goes.  Bare ctypes calls on the solver's own
context: upload alone (returns when the copies are enqueued), upload + stream synchronisation, the solve call with its
result copies, and the whole Python solve(); Khosla on the asymmetric_ksparse shapes (k = 32, 10 objects per person)."""
import ctypes as C, json, sys, time
sys.path.insert(0, ".")
import numpy as np, torch
import sparse_linear_assignment_b200 as S
from sparse_linear_assignment_b200 import generators as G, _lib

lib = _lib.load()
R = 300


def med(xs):
    xs = sorted(xs)
    return round(xs[len(xs) // 2] * 1e6, 1)


for n in (100, 500, 1000, 1900):
    m, k = 10 * n, 32
    rp, c, v = G.kregular_host(n, m, k, seed=1)
    s, z = S.KhoslaSolver.new(n, m, n * k)
    s.load_csr(n, m, rp, c, v)
    hv = s.values()
    for _ in range(20):
        if hv[0] < 0: np.negative(hv, out=hv)
        s._dirty = True
        s.solve(z, False, None)
    ctx = s._context()
    stream = torch.cuda.ExternalStream(s._context_stream())
    a_rp, a_c, a_v = s._i_starts_stops.addr, s._column_indices.addr, s._values.addr
    nnz = n * k
    st = _lib.SlaStats()
    nan = float("nan")
    p2o, o2p, a_p2o, a_o2p = s._outputs(z)
    t_up, t_up_sync, t_solve, t_py, t_up_plain = [], [], [], [], []
    for _ in range(R):
        if hv[0] < 0: np.negative(hv, out=hv)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = lib.sla_upload_csr_negating(ctx, n, m, a_rp, a_c, a_v, nnz, 1)
        t1 = time.perf_counter()
        stream.synchronize()
        t2 = time.perf_counter()
        rc |= lib.sla_khosla_solve(ctx, 0, nan, a_p2o, a_o2p, None, C.byref(st))
        t3 = time.perf_counter()
        assert rc == 0
        t_up.append(t1 - t0); t_up_sync.append(t2 - t0); t_solve.append(t3 - t2)
        # upload without the negation pass (values already negative: nothing to flip)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lib.sla_upload_csr(ctx, n, m, a_rp, a_c, a_v, nnz)
        t_up_plain.append(time.perf_counter() - t0)
        stream.synchronize()
    for _ in range(R):
        if hv[0] < 0: np.negative(hv, out=hv)
        s._dirty = True
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        s.solve(z, False, None)
        t_py.append(time.perf_counter() - t0)
    print(json.dumps({"n": n, "m": m, "k": k, "csr_bytes": int(rp.nbytes + c.nbytes + v.nbytes),
                      "upload_call_us": med(t_up), "upload_plain_call_us": med(t_up_plain), "upload_until_on_device_us": med(t_up_sync),
                      "solve_call_with_downloads_us": med(t_solve), "device_ms_solve_us": round(st.ms_solve * 1e3, 1),
                      "python_solve_us": med(t_py), "kernel_launches": st.kernel_launches, "rounds": st.rounds,
                      "objective": s.get_objective(z), "num_unassigned": int(z.num_unassigned)}), flush=True)
    s.close()
