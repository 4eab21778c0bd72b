"""Batches of independent instances (BASELINE.json config 4) over `sla_batch_*`: one CTA per instance.

The reference has no batch API -- its bench harness clones a solver per problem (benches/benchmark.rs:109,137) -- so
this is a new entry point; each instance is solved exactly as `KhoslaSolver::solve` / `ForwardAuctionSolver::solve`
would solve it alone (own sign normalisation, eps, threshold, eps-scaling phases).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import SlaStats
from .solver import host_array

__all__ = ["BatchSolver"]


class BatchSolver:
    def __init__(self, kind: str, device: int = 0):
        assert kind in ("khosla", "forward")
        self.kind = kind
        self.device = device
        self._ctx = None
        self.n_inst = 0
        self.row_off = self.col_off = None
        self.last_total: Optional[dict] = None

    def _context(self):
        if self._ctx is None:
            lib = _lib.load()
            ctx = C.c_void_p()
            rc = lib.sla_ctx_create(self.device, 1, 1, 1, C.byref(ctx))
            if rc != _lib.SLA_OK:
                msg = lib.sla_last_error(None)
                raise _lib.SlaError(rc, msg.decode() if msg else "")
            self._ctx = ctx
        return self._ctx

    def close(self):
        if self._ctx is not None:
            _lib.load().sla_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, instances: Sequence[Tuple[int, int, np.ndarray, np.ndarray, np.ndarray]]) -> None:
        """instances: (num_rows, num_cols, row_ptr[num_rows+1] (local, starting at 0), cols (local), vals)."""
        n = len(instances)
        row_off = np.zeros(n + 1, dtype=np.uint32)
        col_off = np.zeros(n + 1, dtype=np.uint32)
        rps, cs, vs = [], [], []
        arc = 0
        for b, (nr, nc, rp, c, v) in enumerate(instances):
            row_off[b + 1] = row_off[b] + nr
            col_off[b + 1] = col_off[b] + nc
            rp = np.asarray(rp, dtype=np.int64)
            rps.append(rp[:-1] + arc)
            arc += int(rp[-1])
            cs.append(np.asarray(c, dtype=np.uint32))
            vs.append(np.asarray(v, dtype=np.float64))
        row_ptr = np.concatenate(rps + [np.array([arc])]).astype(np.uint32)
        cols = np.ascontiguousarray(np.concatenate(cs))
        vals = np.ascontiguousarray(np.concatenate(vs))
        ctx = self._context()
        _lib.check(ctx, _lib.load().sla_batch_upload(ctx, n, row_off.ctypes.data, col_off.ctypes.data, row_ptr.ctypes.data,
                                                     cols.ctypes.data, vals.ctypes.data))
        self.n_inst, self.row_off, self.col_off = n, row_off, col_off

    def generate_device(self, n_inst, first_id, num_rows, num_cols, k, seed=0, value_lo=300, value_hi=1000, planted=True):
        """Instance b is generators.kregular_host(num_rows, num_cols, k, seed=seed + first_id + b, planted=...)."""
        ctx = self._context()
        _lib.check(ctx, _lib.load().sla_batch_generate_device(ctx, n_inst, first_id, num_rows, num_cols, k, seed, value_lo,
                                                              value_hi, int(bool(planted))))
        self.n_inst = n_inst
        self.row_off = (np.arange(n_inst + 1, dtype=np.uint64) * num_rows).astype(np.uint32)
        self.col_off = (np.arange(n_inst + 1, dtype=np.uint64) * num_cols).astype(np.uint32)

    def solve(self, maximize=False, eps=None, start_eps=None, max_iterations=None, download=True, per_instance=True):
        """Returns dict(p2o, o2p, prices (concatenated, instance-local indices; None when download=False),
        stats=[per-instance dicts], total=dict)."""
        ctx = self._context()
        nan = float("nan")
        if self.row_off is None:
            raise _lib.SlaError(_lib.SLA_ERR_STATE, "BatchSolver.solve called before upload / generate_device")
        tr, tc = int(self.row_off[-1]), int(self.col_off[-1])
        if download:
            # page-locked result buffers are kept across solves of same-shaped batches (cudaMallocHost costs milliseconds);
            # the returned arrays are therefore overwritten by the next solve(download=True) of this BatchSolver
            if getattr(self, "_out_shape", None) != (tr, tc):
                self._out = (host_array(tr, np.uint32), host_array(tc, np.uint32), host_array(tc, np.float64))
                self._out_shape = (tr, tc)
            p2o, o2p, prices = self._out
        else:
            p2o = o2p = prices = None
        per = (SlaStats * self.n_inst)() if per_instance else None
        total = SlaStats()
        mi = 0 if max_iterations is None else max(int(max_iterations), 1)
        rc = _lib.load().sla_batch_solve(ctx, _lib.ALGO_FORWARD if self.kind == "forward" else _lib.ALGO_KHOSLA,
                                         int(bool(maximize)), nan if eps is None else float(eps),
                                         nan if start_eps is None else float(start_eps), mi,
                                         p2o.ctypes.data if download else None, o2p.ctypes.data if download else None,
                                         prices.ctypes.data if download else None, per, C.byref(total))
        _lib.check(ctx, rc)
        self.last_total = total.as_dict()
        stats: List[dict] = [per[i].as_dict() for i in range(self.n_inst)] if per_instance else []
        return dict(p2o=p2o, o2p=o2p, prices=prices, stats=stats, total=self.last_total)

    def instance_slices(self, b: int):
        return slice(int(self.row_off[b]), int(self.row_off[b + 1])), slice(int(self.col_off[b]), int(self.col_off[b + 1]))
