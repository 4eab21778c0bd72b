"""Host-side mirror of the reference's public API over the C ABI.

The reference crate is Rust (no toolchain in this image), so this module plays the role of the Rust shim for the
tests and benchmarks here: same names, argument meaning, post-conditions and error behaviour as

    AuctionSolver trait            /root/reference/src/solver.rs:8-244
    AuctionSolution                /root/reference/src/solution.rs:22-53
    KhoslaSolver                   /root/reference/src/ksparse.rs:73-260
    ForwardAuctionSolver           /root/reference/src/symmetric.rs:75-508

The host keeps the CSR storage (growable arrays, like the reference's Vecs); `solve` mirrors it into HBM when it
changed and runs the CUDA path through `libsla_b200.so`.  Nothing here computes an auction on the CPU.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import threading
import weakref
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import SlaError, SlaStats

__all__ = ["AuctionSolution", "AuctionSolver", "KhoslaSolver", "ForwardAuctionSolver", "SlaError"]


_U32 = np.dtype(np.uint32)


def _imax(dtype) -> int:
    return int(np.iinfo(dtype).max)


_host_threads_cache = {}


def host_threads() -> int:
    """Host threads one solver may use for the in-place negation: the cores of the box divided by the number of
    co-located ranks (torchrun exports LOCAL_WORLD_SIZE), at most 16.  Evaluated once per value of the two environment
    variables: os.cpu_count() and the environment look-ups cost 15 us, half of the Python time of a small solve."""
    env = os.environ._data if hasattr(os.environ, "_data") else None
    key = (env.get(b"SLA_HOST_THREADS"), env.get(b"LOCAL_WORLD_SIZE")) if env is not None else \
        (os.environ.get("SLA_HOST_THREADS"), os.environ.get("LOCAL_WORLD_SIZE"))
    hit = _host_threads_cache.get(key)
    if hit is not None:
        return hit
    if os.environ.get("SLA_HOST_THREADS"):
        out = max(1, min(16, int(os.environ["SLA_HOST_THREADS"])))
    else:
        cores = os.cpu_count() or 1
        local_world = max(int(os.environ.get("LOCAL_WORLD_SIZE", "1") or 1), 1)
        out = max(1, min(16, cores // local_world))
    _host_threads_cache[key] = out
    return out


_PIN_MIN_BYTES = 1 << 16
_pin_state = {"ok": None}


def host_array(n: int, dtype) -> np.ndarray:
    """An uninitialised host array; large ones are page-locked through the library (sla_host_alloc) so that the
    H2D / D2H copies of the C ABI run at PCIe speed.  Falls back to pageable memory when no device is usable
    (storage only -- nothing is ever computed on the host)."""
    dtype = np.dtype(dtype)
    nbytes = int(n) * dtype.itemsize
    if nbytes >= _PIN_MIN_BYTES and _pin_state["ok"] is not False:
        try:
            lib = _lib.load()
            ptr = C.c_void_p()
            if lib.sla_host_alloc(nbytes, C.byref(ptr)) == _lib.SLA_OK and ptr.value:
                _pin_state["ok"] = True
                arr = np.frombuffer((C.c_char * nbytes).from_address(ptr.value), dtype=dtype, count=int(n))
                weakref.finalize(arr, lib.sla_host_free, ptr.value)
                return arr
            _pin_state["ok"] = False
        except Exception:
            _pin_state["ok"] = False
    return np.empty(int(n), dtype=dtype)


class _Vec:
    """A growable typed array with Vec semantics (len <= capacity, amortised push)."""

    def __init__(self, dtype, capacity: int = 0):
        self.a = host_array(max(int(capacity), 1), dtype)
        self.len = 0

    @property
    def a(self) -> np.ndarray:
        return self._a

    @a.setter
    def a(self, arr: np.ndarray):
        self._a = arr
        self.addr = arr.ctypes.data     # base address for the C ABI, taken once per (re)allocation: ndarray.ctypes costs ~1.5 us

    def reserve(self, n: int):
        if n > self.a.size:
            cap = self.a.size
            while cap < n:
                cap *= 2
            b = host_array(cap, self.a.dtype)
            b[: self.len] = self.a[: self.len]
            self.a = b

    def push(self, x):
        self.reserve(self.len + 1)
        self.a[self.len] = x
        self.len += 1

    def extend(self, xs):
        n = len(xs)
        self.reserve(self.len + n)
        self.a[self.len: self.len + n] = xs
        self.len += n

    def clear(self):
        self.len = 0

    def resize(self, n: int, fill):
        self.reserve(n)
        if n > self.len:
            self.a[self.len: n] = fill
        self.len = n

    def assign(self, arr):
        arr = np.asarray(arr, dtype=self.a.dtype)
        self.reserve(arr.size)
        self.a[: arr.size] = arr
        self.len = arr.size

    @property
    def view(self) -> np.ndarray:
        return self.a[: self.len]


class AuctionSolution:
    """reference src/solution.rs:22-53"""

    def __init__(self, row_capacity: int = 0, column_capacity: int = 0, index_dtype=np.uint32):
        self.index_dtype = np.dtype(index_dtype)
        self.person_to_object = np.empty(0, dtype=self.index_dtype)
        self.object_to_person = np.empty(0, dtype=self.index_dtype)
        self.eps = float("nan")
        self.num_unassigned = _imax(self.index_dtype)

    @classmethod
    def new(cls, row_capacity: int, column_capacity: int, index_dtype=np.uint32) -> "AuctionSolution":
        return cls(row_capacity, column_capacity, index_dtype)

    def clone(self) -> "AuctionSolution":
        z = AuctionSolution(0, 0, self.index_dtype)
        z.person_to_object = self.person_to_object.copy()
        z.object_to_person = self.object_to_person.copy()
        z.eps, z.num_unassigned = self.eps, self.num_unassigned
        return z


def _ensure(cond: bool, what: str):
    if not cond:
        raise SlaError(_lib.SLA_ERR_INVALID, what)


class AuctionSolver:
    """Shared trait with default methods (reference src/solver.rs:8-244)."""

    _ALGO = None

    def __init__(self, row_capacity: int, column_capacity: int, arcs_capacity: int, index_dtype=np.uint32,
                 device: int = 0):
        self.index_dtype = np.dtype(index_dtype)
        assert self.index_dtype in (np.dtype(np.uint16), np.dtype(np.uint32)), "UnsignedInt is u16 or u32"
        self.imax = _imax(self.index_dtype)
        self.device = device
        self._caps = (int(row_capacity), int(column_capacity), int(arcs_capacity))
        self._num_rows = 0
        self._num_cols = 0
        self._i_starts_stops = _Vec(self.index_dtype, row_capacity + 1)
        self._j_counts = _Vec(self.index_dtype, row_capacity)
        self._prices = _Vec(np.float64, column_capacity)
        self._column_indices = _Vec(self.index_dtype, arcs_capacity)
        self._values = _Vec(np.float64, arcs_capacity)
        self.nits = 0
        self._last_stats: Optional[dict] = None
        self._last_struct: Optional[SlaStats] = None
        self._out_cache = None      # (person_to_object, object_to_person, their base addresses) of the last solve
        self._ctx = None
        self._dirty = True
        self._device_only = False   # CSR was generated in HBM (generators.kregular_device): no host copy exists
        self._prices_on_device = False

    @property
    def last_stats(self) -> Optional[dict]:
        """sla_stats of the last solve as a dict (built on first access: 7 us that a small solve need not pay)."""
        if self._last_stats is None and self._last_struct is not None:
            self._last_stats = self._last_struct.as_dict()
        return self._last_stats

    @last_stats.setter
    def last_stats(self, value: Optional[dict]):
        self._last_stats, self._last_struct = value, None

    # ---- construction ------------------------------------------------------------------------------------
    @classmethod
    def new(cls, row_capacity: int, column_capacity: int, arcs_capacity: int, index_dtype=np.uint32, device: int = 0
            ) -> Tuple["AuctionSolver", AuctionSolution]:
        """solver.rs:9-13: returns (solver, solution)."""
        return (cls(row_capacity, column_capacity, arcs_capacity, index_dtype, device),
                AuctionSolution(row_capacity, column_capacity, index_dtype))

    def clone(self):
        """#[derive(Clone)]: deep copy of the host state; the device context is re-created lazily."""
        other = type(self)(*self._caps, index_dtype=self.index_dtype, device=self.device)
        other._num_rows, other._num_cols = self._num_rows, self._num_cols
        self.prices()    # pull resident prices into the host copy before it is duplicated
        for name in ("_i_starts_stops", "_j_counts", "_prices", "_column_indices", "_values"):
            getattr(other, name).assign(getattr(self, name).view)
        for name in ("nits", "nreductions", "optimal_soln_found", "max_iterations"):
            if hasattr(self, name):
                setattr(other, name, getattr(self, name))
        return other

    def close(self):
        if self._ctx is not None:
            _lib.load().sla_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- accessors (solver.rs:22-38); the *_mut twins mark the device mirror stale -----------------------
    def num_rows(self) -> int:
        return self._num_rows

    def num_cols(self) -> int:
        return self._num_cols

    def prices(self) -> np.ndarray:
        """Final prices of the last solve; fetched from HBM on first access (they stay resident otherwise)."""
        if self._prices_on_device:
            ctx = self._context()
            self._prices.resize(self._num_cols, 0.0)
            _lib.check(ctx, _lib.load().sla_download_solution(ctx, None, None, self._prices.view.ctypes.data))
            self._prices_on_device = False
        return self._prices.view

    def i_starts_stops(self) -> np.ndarray:
        return self._i_starts_stops.view

    def j_counts(self) -> np.ndarray:
        return self._j_counts.view

    def column_indices(self) -> np.ndarray:
        return self._column_indices.view

    def values(self) -> np.ndarray:
        return self._values.view

    def prices_mut(self) -> np.ndarray:
        return self.prices()

    def i_starts_stops_mut(self) -> np.ndarray:
        self._dirty = True
        return self._i_starts_stops.view

    def j_counts_mut(self) -> np.ndarray:
        self._dirty = True
        return self._j_counts.view

    def column_indices_mut(self) -> np.ndarray:
        self._dirty = True
        return self._column_indices.view

    def values_mut(self) -> np.ndarray:
        self._dirty = True
        return self._values.view

    # ---- CSR builder ---------------------------------------------------------------------------------------
    def init(self, num_rows: int, num_cols: int) -> None:
        """solver.rs:191-205"""
        _ensure(num_rows <= num_cols, "num_rows <= num_cols")
        _ensure(num_rows < self.imax, "num_rows < I::MAX")
        self._num_rows, self._num_cols = int(num_rows), int(num_cols)
        self._i_starts_stops.clear()
        self._i_starts_stops.resize(2, 0)
        self._j_counts.clear()
        self._j_counts.push(0)
        self._column_indices.clear()
        self._values.clear()
        self._dirty = True
        self._device_only = False

    def add_value(self, row: int, column: int, value: float) -> None:
        """solver.rs:41-66"""
        current_row = self._j_counts.len - 1
        _ensure(row == current_row or row == current_row + 1, "rows must arrive in non-decreasing order")
        prev = int(self._i_starts_stops.a[current_row + 1])
        _ensure(prev + 1 <= self.imax, "i_starts_stops vector is longer then max value of type")
        if row > current_row:
            _ensure(int(self._j_counts.a[current_row]) > 0, "previous row is empty")
            self._i_starts_stops.push(prev + 1)
            self._j_counts.push(1)
        else:
            self._i_starts_stops.a[current_row + 1] = prev + 1
            self._j_counts.a[current_row] += 1
        self._column_indices.push(column)
        self._values.push(value)
        self._dirty = True

    def extend_from_values(self, row: int, columns, values) -> None:
        """solver.rs:69-101"""
        columns = np.asarray(columns)
        values = np.asarray(values, dtype=np.float64)
        _ensure(columns.size == values.size, "columns.len() == values.len()")
        current_row = self._j_counts.len - 1
        _ensure(row == current_row or row == current_row + 1, "rows must arrive in non-decreasing order")
        inc = int(columns.size)
        _ensure(inc <= self.imax, "columns slice is longer then max value of type")
        prev = int(self._i_starts_stops.a[current_row + 1])
        _ensure(prev + inc <= self.imax, "i_starts_stops vector is longer then max value of type")
        if row > current_row:
            _ensure(int(self._j_counts.a[current_row]) > 0, "previous row is empty")
            self._i_starts_stops.push(prev + inc)
            self._j_counts.push(inc)
        else:
            self._i_starts_stops.a[current_row + 1] = prev + inc
            self._j_counts.a[current_row] += inc
        self._column_indices.extend(columns.astype(self.index_dtype, copy=False))
        self._values.extend(values)
        self._dirty = True

    def load_csr(self, num_rows: int, num_cols: int, row_ptr, cols, vals) -> None:
        """Bulk equivalent of init + one extend_from_values per row (not in the reference; same resulting state)."""
        self.init(num_rows, num_cols)
        row_ptr = np.asarray(row_ptr)
        _ensure(row_ptr.size == num_rows + 1 and int(row_ptr[0]) == 0, "row_ptr must have num_rows + 1 entries")
        counts = np.diff(row_ptr.astype(np.int64))
        _ensure(bool(np.all(counts[:-1] > 0)) if num_rows > 1 else True, "previous row is empty")
        _ensure(int(row_ptr[-1]) <= self.imax, "i_starts_stops vector is longer then max value of type")
        self._i_starts_stops.assign(row_ptr)
        self._j_counts.assign(counts)
        self._column_indices.assign(cols)
        self._values.assign(vals)
        self._dirty = True

    def num_of_arcs(self) -> int:
        """solver.rs:104-106"""
        return self._column_indices.len

    def validate_input(self) -> None:
        """solver.rs:232-243"""
        arcs = self.num_of_arcs()
        _ensure(arcs > 0, "arcs_count > 0")
        _ensure(self._num_rows > 0 and self._num_cols > 0, "num_rows > 0 && num_cols > 0")
        _ensure(arcs < self.imax, "arcs_count < I::MAX")
        _ensure(arcs == self._column_indices.len == self._values.len, "column_indices.len() == values.len()")

    # ---- post-processing (host side, like the reference) ---------------------------------------------------
    def get_objective(self, solution: AuctionSolution) -> float:
        """solver.rs:110-142.  Left-to-right accumulation in row order, so the result is bit-identical to the
        reference's loop for any weights (np.cumsum accumulates sequentially)."""
        nnz = self.num_of_arcs()
        vals = self._values.view
        positive = (vals[0] if nnz else 0.0) >= 0.0
        n = self._num_rows
        counts = self._j_counts.view[:n].astype(np.int64)
        p2o = np.asarray(solution.person_to_object)[:n]
        chosen = np.repeat(p2o.astype(np.int64), counts)
        assigned = np.repeat(p2o != self.imax, counts)
        hit = (self._column_indices.view.astype(np.int64) == chosen) & assigned
        terms = vals[hit]
        if terms.size == 0:
            return 0.0
        if not positive:
            terms = -terms
        return float(np.cumsum(terms)[-1])

    def get_toleration(self, max_abs_cost: float) -> float:
        """solver.rs:144-146"""
        l = math.log2(max_abs_cost + 1e-7) if max_abs_cost + 1e-7 > 0 else float("-inf")
        li = 0 if not (l > 0.0) else min(int(l), 0xFFFFFFFF)
        e = (53 - li) & 0xFFFFFFFF
        return 1.0 / float(1 << e) if e < 64 else float("inf")

    def ecs_satisfied(self, person_to_object, eps: float, toleration: float) -> bool:
        """solver.rs:154-189 on the host copies (prices() holds the final prices after a solve)."""
        n = self._num_rows
        counts = self._j_counts.view[:n].astype(np.int64)
        starts = self._i_starts_stops.view[:n].astype(np.int64)
        p2o = np.asarray(person_to_object)[:n].astype(np.int64)
        cols = self._column_indices.view.astype(np.int64)
        vals = self._values.view
        prices = self.prices()
        row_of = np.repeat(np.arange(n), counts)
        match = cols == p2o[row_of]
        chosen = np.full(n, -np.inf)
        idx = np.nonzero(match)[0]
        chosen[row_of[idx]] = vals[idx]          # later arcs overwrite earlier ones: the LAST match wins
        lhs = chosen - prices[p2o] + toleration
        del starts
        return not bool(np.any(lhs[row_of] < vals - prices[cols] - eps))

    def init_solve(self, solution: AuctionSolution, maximize: bool) -> None:
        """solver.rs:207-230 on the host copies (the device path applies the same normalisation in HBM)."""
        vals = self._values.view
        positive = (vals[0] if vals.size else 0.0) >= 0.0
        if bool(maximize) ^ bool(positive):
            vals *= -1.0
            self._dirty = True
        self._prices_on_device = False
        self._prices.clear()
        self._prices.resize(self._num_cols, 0.0)
        solution.person_to_object = np.full(self._num_rows, self.imax, dtype=self.index_dtype)
        solution.object_to_person = np.full(self._num_cols, self.imax, dtype=self.index_dtype)
        solution.num_unassigned = self._num_rows

    # ---- device plumbing -------------------------------------------------------------------------------------
    def _context(self):
        if self._ctx is None:
            lib = _lib.load()
            ctx = C.c_void_p()
            rc = lib.sla_ctx_create(self.device, self._caps[0], self._caps[1], self._caps[2], C.byref(ctx))
            if rc != _lib.SLA_OK:
                msg = lib.sla_last_error(None)
                raise SlaError(rc, msg.decode() if msg else "")
            self._ctx = ctx
            self._dirty = True
        return self._ctx

    def _context_stream(self) -> int:
        """The context's cudaStream_t as an integer (for torch.cuda.ExternalStream / event timing)."""
        return int(_lib.load().sla_ctx_stream(self._context()) or 0)

    def last_upload(self) -> Tuple[int, int]:
        """(bytes the last upload moved host -> device, bytes per value on the wire: 2, 4 or 8)."""
        ctx = self._context()
        b, w = C.c_uint64(), C.c_uint32()
        _lib.check(ctx, _lib.load().sla_last_upload(ctx, C.byref(b), C.byref(w)))
        return int(b.value), int(w.value)

    def scan_value_bytes(self) -> int:
        """Bytes per value the grid-wide uniform-degree scans read on the resident CSR (sla_scan_value_bytes): 2 after a
        u16 upload, else 8."""
        ctx = self._context()
        w = C.c_uint32()
        _lib.check(ctx, _lib.load().sla_scan_value_bytes(ctx, C.byref(w)))
        return int(w.value)

    def set_option(self, key: str, value: int) -> None:
        ctx = self._context()
        _lib.check(ctx, _lib.load().sla_set_option(ctx, key.encode(), int(value)))

    def _sync_device(self, maximize=None):
        """Mirrors the host CSR into HBM when it changed.  With `maximize` given (the solve paths), a pending in-place
        sign normalisation of `values` (solver.rs:209-216) is started on the library's worker threads as soon as the
        values have crossed PCIe (sla_upload_csr_negating), overlapping the rest of the upload and the solve."""
        ctx = self._context()
        if self._dirty:
            n, nnz = self._num_rows, self.num_of_arcs()
            _ensure(self._i_starts_stops.len >= n + 1, "fewer rows populated than num_rows")
            if self.index_dtype == _U32:        # the host vectors are what the C ABI reads: no copies, cached addresses
                row_ptr = cols = None
                rp_addr, cols_addr = self._i_starts_stops.addr, self._column_indices.addr
            else:                               # u16 indices are widened for the device (kept alive until the call returns)
                row_ptr = np.ascontiguousarray(self._i_starts_stops.view[: n + 1], dtype=np.uint32)
                cols = np.ascontiguousarray(self._column_indices.view, dtype=np.uint32)
                rp_addr, cols_addr = row_ptr.ctypes.data, cols.ctypes.data
            vals_addr = self._values.addr
            flip = maximize is not None and (bool(maximize) ^ bool((self._values.a[0] if nnz else 0.0) >= 0.0))
            # any size: small instances are negated in the library's single staging pass, large ones by its worker pool
            self._pre_negated = bool(flip)
            if self._pre_negated:
                rc = _lib.load().sla_upload_csr_negating(ctx, n, self._num_cols, rp_addr, cols_addr, vals_addr, nnz,
                                                         host_threads())
            else:
                rc = _lib.load().sla_upload_csr(ctx, n, self._num_cols, rp_addr, cols_addr, vals_addr, nnz)
            del row_ptr, cols
            _lib.check(ctx, rc)
            self._dirty = False
        return ctx

    def _begin_negation(self, maximize: bool):
        """solver.rs:209-216 on the host copy: decided by the sign of the first value; runs on helper threads while
        the GPU solves (the device applies the same sign on the fly).  Returns (thread or None, flip)."""
        if self._device_only:
            return None, None
        if getattr(self, "_pre_negated", False):
            # the upload of this very solve already started the negation on the library's worker threads
            # (sla_upload_csr_negating); they are joined inside the solve call
            return None, True
        vals = self._values.view
        flip = bool(maximize) ^ bool((vals[0] if vals.size else 0.0) >= 0.0)
        if not flip:
            return None, False
        lib = _lib.load()
        if vals.size < (1 << 16):      # too small to be worth a helper thread: negate right here
            lib.sla_host_negate_f64(self._values.addr, vals.size, 1)
            return None, True
        th = threading.Thread(target=lib.sla_host_negate_f64,
                              args=(vals.ctypes.data, vals.size, host_threads()))
        th.start()
        return th, True

    def _check_solve(self, ctx, rc):
        """Raises on a failed solve.  A failure (timeout_s, CUDA error, safety round limit) can come after the device
        has already applied the sign normalisation and the upload's workers have negated the host copy, so both sides
        are consistent; only the marker that the *next* solve's normalisation was pre-applied must not survive, or
        every later solve would expect a flip the device no longer reports."""
        try:
            _lib.check(ctx, rc)
        except SlaError:
            self._pre_negated = False
            raise

    def _finish(self, solution: AuctionSolution, stats: SlaStats, p2o, o2p, negation=(None, None)):
        th, flip = negation
        if th is not None:
            th.join()
        if flip is not None and bool(stats.values_negated) != flip:
            raise SlaError(_lib.SLA_ERR_STATE, "host / device disagree on the sign normalisation of values")
        self._pre_negated = False
        if self.index_dtype == np.dtype(np.uint32):
            solution.person_to_object, solution.object_to_person = p2o, o2p
        else:   # SLA_NONE truncates to u16::MAX
            solution.person_to_object = p2o.astype(self.index_dtype)
            solution.object_to_person = o2p.astype(self.index_dtype)
        solution.num_unassigned = int(stats.num_unassigned)
        solution.eps = float(stats.eps)
        self.nits = int(stats.nits)
        self._last_stats, self._last_struct = None, stats

    def _outputs(self, solution: AuctionSolution):
        """Output buffers and their base addresses: the caller's solution vectors are reused when they already have the
        right shape (the reference resizes the caller's Vecs in place, solver.rs:221-228); prices stay in HBM until asked
        for.  The vectors a previous solve of this solver handed out are recognised by identity (the cache holds a
        reference, so numpy refuses to resize them in place)."""
        oc = self._out_cache
        if oc is not None and solution.person_to_object is oc[0] and solution.object_to_person is oc[1] \
                and oc[0].size == self._num_rows and oc[1].size == self._num_cols:
            return oc

        def fit(arr, n):
            if isinstance(arr, np.ndarray) and arr.dtype == np.uint32 and arr.size == n and arr.flags.c_contiguous \
                    and arr.flags.writeable:
                return arr
            return host_array(n, np.uint32)
        p2o = fit(solution.person_to_object, self._num_rows)
        o2p = fit(solution.object_to_person, self._num_cols)
        self._out_cache = (p2o, o2p, p2o.ctypes.data, o2p.ctypes.data)
        return self._out_cache

    def device_objective(self) -> float:
        """sla_get_objective on the resident solution (exact for integer-valued weights)."""
        ctx = self._context()
        out = C.c_double()
        _lib.check(ctx, _lib.load().sla_get_objective(ctx, C.byref(out)))
        return out.value

    def device_ecs_satisfied(self, eps: float, toleration: float) -> bool:
        ctx = self._context()
        out = C.c_int()
        _lib.check(ctx, _lib.load().sla_ecs_satisfied(ctx, eps, toleration, C.byref(out)))
        return bool(out.value)

    def device_validate_matching(self):
        ctx = self._context()
        un, ok = C.c_uint32(), C.c_int()
        _lib.check(ctx, _lib.load().sla_validate_matching(ctx, C.byref(un), C.byref(ok)))
        return un.value, bool(ok.value)

    def round_profile(self):
        ctx = self._context()
        lib = _lib.load()
        n = C.c_size_t()
        _lib.check(ctx, lib.sla_get_round_profile(ctx, None, 0, C.byref(n)))
        buf = (_lib.SlaRoundProfile * max(n.value, 1))()
        _lib.check(ctx, lib.sla_get_round_profile(ctx, buf, n.value, C.byref(n)))
        return [buf[i].as_dict() for i in range(n.value)]

    def solve(self, solution: AuctionSolution, maximize: bool, eps: Optional[float] = None) -> None:
        raise NotImplementedError

    def solve_resident(self, maximize: bool, eps: Optional[float] = None, **kw) -> dict:
        """Runs the same solve but leaves person_to_object / object_to_person / prices in HBM (no D2H); used by
        bench.py for the device-resident throughput.  Returns the stats dict."""
        if not self._device_only:
            self.validate_input()
        ctx = self._sync_device()
        st = SlaStats()
        nan = float("nan")
        lib = _lib.load()
        if isinstance(self, ForwardAuctionSolver):
            mi = kw.get("max_iterations")
            rc = lib.sla_forward_solve(ctx, int(bool(maximize)), nan if eps is None else float(eps),
                                       nan if kw.get("start_eps") is None else float(kw["start_eps"]),
                                       self.MAX_ITERATIONS if mi is None else max(int(mi), 1), None, None, None,
                                       C.byref(st))
        else:
            rc = lib.sla_khosla_solve(ctx, int(bool(maximize)), nan if eps is None else float(eps), None, None, None,
                                      C.byref(st))
        _lib.check(ctx, rc)
        if st.values_negated and not self._device_only:
            _lib.load().sla_host_negate_f64(self._values.view.ctypes.data, self._values.view.size, 8)
        self._prices_on_device = True
        self.nits = int(st.nits)
        self.last_stats = st.as_dict()
        return self.last_stats

    def download_solution(self, solution: AuctionSolution) -> None:
        ctx = self._context()
        p2o, o2p, p2o_addr, o2p_addr = self._outputs(solution)
        _lib.check(ctx, _lib.load().sla_download_solution(ctx, p2o_addr, o2p_addr, None))
        st = SlaStats(**{k: v for k, v in (self.last_stats or {}).items()})
        self._finish(solution, st, p2o, o2p)


class KhoslaSolver(AuctionSolver):
    """reference src/ksparse.rs:73-260; the bidding loop runs as synchronous Jacobi rounds on the GPU."""

    def solve(self, solution: AuctionSolution, maximize: bool, eps: Optional[float] = None) -> None:
        """ksparse.rs:153-251"""
        if not self._device_only:
            self.validate_input()
        ctx = self._sync_device(maximize)
        p2o, o2p, p2o_addr, o2p_addr = self._outputs(solution)
        st = SlaStats()
        neg = self._begin_negation(maximize)
        rc = _lib.load().sla_khosla_solve(ctx, int(bool(maximize)), float("nan") if eps is None else float(eps),
                                          p2o_addr, o2p_addr, None, C.byref(st))
        if neg[0] is not None:
            neg[0].join()
        self._check_solve(ctx, rc)
        self._prices_on_device = True
        self._finish(solution, st, p2o, o2p, neg)


class ForwardAuctionSolver(AuctionSolver):
    """reference src/symmetric.rs:75-508"""

    REDUCTION_FACTOR = 0.15      # symmetric.rs:189
    MAX_ITERATIONS = 100000      # symmetric.rs:190

    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.max_iterations = self.MAX_ITERATIONS
        self.nreductions = 0
        self.optimal_soln_found = False

    def solve(self, solution: AuctionSolution, maximize: bool, eps: Optional[float] = None) -> None:
        """symmetric.rs:177-185"""
        self.solve_with_params(solution, maximize, eps, None, None)

    def solve_with_params(self, solution: AuctionSolution, maximize: bool, eps: Optional[float] = None,
                          start_eps: Optional[float] = None, max_iterations: Optional[int] = None) -> None:
        """symmetric.rs:217-332"""
        if not self._device_only:
            self.validate_input()
        ctx = self._sync_device(maximize)
        p2o, o2p, p2o_addr, o2p_addr = self._outputs(solution)
        # Some(0) behaves like Some(1) in the reference (the check runs after the first round, symmetric.rs:326)
        self.max_iterations = max(int(max_iterations), 1) if max_iterations is not None else self.MAX_ITERATIONS
        st = SlaStats()
        nan = float("nan")
        neg = self._begin_negation(maximize)
        rc = _lib.load().sla_forward_solve(ctx, int(bool(maximize)), nan if eps is None else float(eps),
                                           nan if start_eps is None else float(start_eps), self.max_iterations,
                                           p2o_addr, o2p_addr, None, C.byref(st))
        if neg[0] is not None:
            neg[0].join()
        self._check_solve(ctx, rc)
        self._prices_on_device = True
        self._finish(solution, st, p2o, o2p, neg)
        self.nreductions = int(st.nreductions)
        self.optimal_soln_found = bool(st.optimal_soln_found)
