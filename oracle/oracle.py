"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

`OracleSolver` mirrors the reference's solver objects (KhoslaSolver / ForwardAuctionSolver sharing the
AuctionSolver trait, /root/reference/src/solver.rs:8-244) over `liboracle.so`; `fixture_ksparse` regenerates
the reference's random test fixtures (solver.rs:261-292); `jacobi_model` runs the CPU model of the device
algorithm (oracle/jacobi_model.c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_SOURCES = ("sla_oracle.c", "fixture_rng.c", "jacobi_model.c", "Makefile")

U32_MAX = 0xFFFFFFFF
U16_MAX = 0xFFFF
KHOSLA, FORWARD = 0, 1


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile when missing or stale."""
    stale = force or not os.path.exists(_LIB_PATH)
    if not stale:
        t = os.path.getmtime(_LIB_PATH)
        stale = any(os.path.getmtime(os.path.join(_HERE, s)) > t for s in _SOURCES)
    if stale:
        subprocess.run(["make", "-s", "-C", _HERE, "-B", "liboracle.so"], check=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _declare(_lib)
    return _lib


_u32p = C.POINTER(C.c_uint32)
_f64p = C.POINTER(C.c_double)


class JmStats(C.Structure):
    _fields_ = [
        ("num_unassigned", C.c_uint32), ("nits", C.c_uint32), ("nreductions", C.c_uint32),
        ("optimal_soln_found", C.c_uint32), ("eps", C.c_double), ("rounds", C.c_uint64), ("bids", C.c_uint64),
        ("bid_arcs", C.c_uint64), ("dropped", C.c_uint32), ("values_negated", C.c_uint32),
        ("restarts", C.c_uint32), ("reserved_", C.c_uint32),
    ]


def _declare(l: C.CDLL) -> None:
    vp = C.c_void_p
    l.orc_solver_new.restype = vp
    l.orc_solver_new.argtypes = [C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t]
    l.orc_solver_free.argtypes = [vp]
    l.orc_solution_new.restype = vp
    l.orc_solution_new.argtypes = [C.c_uint32]
    l.orc_solution_free.argtypes = [vp]
    l.orc_init.argtypes = [vp, C.c_uint32, C.c_uint32]
    l.orc_add_value.argtypes = [vp, C.c_uint32, C.c_uint32, C.c_double]
    l.orc_extend_from_values.argtypes = [vp, C.c_uint32, _u32p, C.c_size_t, _f64p, C.c_size_t]
    l.orc_load_csr.argtypes = [vp, C.c_uint32, C.c_uint32, _u32p, _u32p, _f64p]
    l.orc_num_of_arcs.restype = C.c_size_t
    l.orc_num_of_arcs.argtypes = [vp]
    l.orc_validate_input.argtypes = [vp]
    l.orc_get_objective.restype = C.c_double
    l.orc_get_objective.argtypes = [vp, vp]
    l.orc_get_toleration.restype = C.c_double
    l.orc_get_toleration.argtypes = [C.c_double]
    l.orc_ecs_satisfied.argtypes = [vp, _u32p, C.c_double, C.c_double]
    l.orc_khosla_solve.argtypes = [vp, vp, C.c_int, C.c_double]
    l.orc_forward_solve.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double, C.c_uint32]
    l.orc_push_all_left.argtypes = [_u32p, C.c_size_t, _u32p, C.c_uint32, C.c_uint32, C.c_uint32]
    for name in ("orc_num_rows", "orc_num_cols", "orc_nits", "orc_nreductions"):
        getattr(l, name).restype = C.c_uint32
        getattr(l, name).argtypes = [vp]
    l.orc_optimal_soln_found.argtypes = [vp]
    l.orc_bid_arcs.restype = C.c_uint64
    l.orc_bid_arcs.argtypes = [vp]
    for name, rt in (("orc_prices", _f64p), ("orc_values", _f64p), ("orc_column_indices", _u32p),
                     ("orc_i_starts_stops", _u32p), ("orc_j_counts", _u32p), ("orc_person_to_object", _u32p),
                     ("orc_object_to_person", _u32p)):
        getattr(l, name).restype = rt
        getattr(l, name).argtypes = [vp]
    for name in ("orc_prices_len", "orc_i_starts_stops_len", "orc_j_counts_len", "orc_person_to_object_len",
                 "orc_object_to_person_len"):
        getattr(l, name).restype = C.c_size_t
        getattr(l, name).argtypes = [vp]
    l.orc_num_unassigned.restype = C.c_uint32
    l.orc_num_unassigned.argtypes = [vp]
    l.orc_solution_eps.restype = C.c_double
    l.orc_solution_eps.argtypes = [vp]
    l.orc_fixture_ksparse.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_double,
                                      _u32p, _u32p, _f64p]
    l.orc_chacha8_u64_stream.argtypes = [C.c_uint64, C.POINTER(C.c_uint64), C.c_size_t]
    l.jm_pack_bid.restype = C.c_uint64
    l.jm_pack_bid.argtypes = [C.c_double, C.c_uint32, C.c_uint32]
    l.jm_set_khosla_scaling.argtypes = [C.c_int]
    l.jm_set_khosla_scaling.restype = None
    l.jm_solve.argtypes = [C.c_int, C.c_uint32, C.c_uint32, _u32p, _u32p, _f64p, C.c_int, C.c_double, C.c_double,
                           C.c_uint32, _u32p, _u32p, _f64p, C.POINTER(JmStats)]


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p32(a):
    return a.ctypes.data_as(_u32p)


def _p64(a):
    return a.ctypes.data_as(_f64p)


class OracleError(RuntimeError):
    """Raised where the reference returns Err (anyhow ensure!/anyhow!)."""


class OracleSolver:
    """One solver + one solution, reference semantics.  kind: 'khosla' | 'forward'; imax: I::MAX (u32 or u16)."""

    def __init__(self, kind: str, row_cap: int = 0, col_cap: int = 0, arc_cap: int = 0, imax: int = U32_MAX):
        assert kind in ("khosla", "forward")
        self.kind = kind
        self.imax = imax
        self._l = lib()
        self._s = self._l.orc_solver_new(imax, row_cap, col_cap, arc_cap)
        self._z = self._l.orc_solution_new(imax)

    def __del__(self):
        try:
            self._l.orc_solver_free(self._s)
            self._l.orc_solution_free(self._z)
        except Exception:
            pass

    def _chk(self, rc):
        if rc != 0:
            raise OracleError("reference would return Err")

    def init(self, num_rows, num_cols):
        self._chk(self._l.orc_init(self._s, num_rows, num_cols))

    def add_value(self, row, column, value):
        self._chk(self._l.orc_add_value(self._s, row, column, float(value)))

    def extend_from_values(self, row, columns, values):
        c, v = _u32(columns), _f64(values)
        self._chk(self._l.orc_extend_from_values(self._s, row, _p32(c), c.size, _p64(v), v.size))

    def load_csr(self, num_rows, num_cols, row_ptr, cols, vals):
        r, c, v = _u32(row_ptr), _u32(cols), _f64(vals)
        self._chk(self._l.orc_load_csr(self._s, num_rows, num_cols, _p32(r), _p32(c), _p64(v)))

    def load_dense(self, costs):
        costs = np.asarray(costs, dtype=np.float64)
        n, m = costs.shape
        self.init(n, m)
        for i in range(n):
            self.extend_from_values(i, np.arange(m), costs[i])

    def solve(self, maximize=False, eps=None, start_eps=None, max_iterations=None):
        nan = float("nan")
        if self.kind == "khosla":
            rc = self._l.orc_khosla_solve(self._s, self._z, int(maximize), nan if eps is None else eps)
        else:
            rc = self._l.orc_forward_solve(self._s, self._z, int(maximize), nan if eps is None else eps,
                                           nan if start_eps is None else start_eps, max_iterations or 0)
        self._chk(rc)

    def get_objective(self):
        return self._l.orc_get_objective(self._s, self._z)

    def ecs_satisfied(self, eps, toleration):
        p = self.person_to_object
        return bool(self._l.orc_ecs_satisfied(self._s, _p32(p), eps, toleration))

    def num_of_arcs(self):
        return self._l.orc_num_of_arcs(self._s)

    def _arr(self, ptr, n, dtype):
        if n == 0:
            return np.zeros(0, dtype=dtype)
        return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)

    @property
    def person_to_object(self):
        return self._arr(self._l.orc_person_to_object(self._z), self._l.orc_person_to_object_len(self._z), np.uint32)

    @property
    def object_to_person(self):
        return self._arr(self._l.orc_object_to_person(self._z), self._l.orc_object_to_person_len(self._z), np.uint32)

    @property
    def prices(self):
        return self._arr(self._l.orc_prices(self._s), self._l.orc_prices_len(self._s), np.float64)

    @property
    def values(self):
        return self._arr(self._l.orc_values(self._s), self.num_of_arcs(), np.float64)

    @property
    def i_starts_stops(self):
        return self._arr(self._l.orc_i_starts_stops(self._s), self._l.orc_i_starts_stops_len(self._s), np.uint32)

    @property
    def j_counts(self):
        return self._arr(self._l.orc_j_counts(self._s), self._l.orc_j_counts_len(self._s), np.uint32)

    num_unassigned = property(lambda self: self._l.orc_num_unassigned(self._z))
    eps = property(lambda self: self._l.orc_solution_eps(self._z))
    nits = property(lambda self: self._l.orc_nits(self._s))
    nreductions = property(lambda self: self._l.orc_nreductions(self._s))
    optimal_soln_found = property(lambda self: bool(self._l.orc_optimal_soln_found(self._s)))
    bid_arcs = property(lambda self: self._l.orc_bid_arcs(self._s))


def get_toleration(c: float) -> float:
    return lib().orc_get_toleration(c)


def push_all_left(data, mapper, num_ints, size, imax=U32_MAX):
    d, m = _u32(data).copy(), _u32(mapper).copy()
    lib().orc_push_all_left(_p32(d), d.size, _p32(m), num_ints, size, imax)
    return d, m


def fixture_ksparse(num_rows, num_cols, k, max_value, val_seed=1, filter_seed=2):
    """populate_with_ksparse_input (solver.rs:261-292) -> (row_ptr, cols, vals)."""
    row_ptr = np.zeros(num_rows + 1, dtype=np.uint32)
    cols = np.zeros(num_rows * k, dtype=np.uint32)
    vals = np.zeros(num_rows * k, dtype=np.float64)
    lib().orc_fixture_ksparse(val_seed, filter_seed, num_rows, num_cols, k, max_value, _p32(row_ptr), _p32(cols),
                              _p64(vals))
    return row_ptr, cols, vals


def chacha8_u64_stream(seed, n):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_chacha8_u64_stream(seed, out.ctypes.data_as(C.POINTER(C.c_uint64)), n)
    return out


def jacobi_model(algo, num_rows, num_cols, row_ptr, cols, vals, maximize=False, eps=None, start_eps=None,
                 max_iterations=None, khosla_scaling=True):
    """CPU model of the device algorithm.  Returns dict(p2o, o2p, prices, stats).  khosla_scaling mirrors the library
    option of the same name (Khosla rounds under an eps-schedule on square instances)."""
    lib().jm_set_khosla_scaling(int(bool(khosla_scaling)))
    r, c, v = _u32(row_ptr), _u32(cols), _f64(vals)
    p2o = np.zeros(num_rows, dtype=np.uint32)
    o2p = np.zeros(num_cols, dtype=np.uint32)
    prices = np.zeros(num_cols, dtype=np.float64)
    st = JmStats()
    nan = float("nan")
    lib().jm_solve(KHOSLA if algo in (KHOSLA, "khosla") else FORWARD, num_rows, num_cols, _p32(r), _p32(c), _p64(v),
                   int(maximize), nan if eps is None else eps, nan if start_eps is None else start_eps,
                   max_iterations or 0, _p32(p2o), _p32(o2p), _p64(prices), C.byref(st))
    stats = {f: getattr(st, f) for f, _ in JmStats._fields_}
    return dict(p2o=p2o, o2p=o2p, prices=prices, stats=stats)


def pack_bid(bid, person, pbits):
    return lib().jm_pack_bid(bid, person, pbits)
