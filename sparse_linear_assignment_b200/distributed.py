"""Row-partitioned KhoslaSolver across ranks (BASELINE.json config 5): one process per GPU, `torch.distributed`
for the plumbing (NCCL over NVLink on GPUs, gloo in the CPU tests).

Every rank owns a contiguous block of persons (CSR rows) and a replica of the object state.  One synchronous round:

    engine.bid()                               local bid scan, local maxima of the packed (bid key, person) words
    all_reduce(words, MAX)                     int64 view of the uint64 words (bit 63 is always clear)
    engine.claim()                             local winners publish their exact f64 bid, losers re-queue
    all_reduce(candidates, MAX)                f64, -inf = no bid
    engine.assign()                            every rank applies all winners to its replica
    all_reduce(local queue length, SUM) == 0   -> done

The engine is `CudaShardEngine` (the C ABI `sla_part_*`) in production; the CPU tests drive the same loop with a
model engine (tests/test_distributed_gloo.py), which is how the host-side logic is covered without GPUs.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch
import torch.distributed as dist

from . import _lib
from ._lib import SlaStats

__all__ = ["PartitionedKhoslaSolver", "CudaShardEngine", "shard_rows"]


def shard_rows(global_rows: int, world: int, rank: int):
    """Contiguous, balanced row ranges: rank r owns [begin, begin + count)."""
    base, extra = divmod(global_rows, world)
    begin = rank * base + min(rank, extra)
    return begin, base + (1 if rank < extra else 0)


class _DevBuf:
    """Exposes a raw device pointer through __cuda_array_interface__ so that torch can alias it (zero copy)."""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class CudaShardEngine:
    """This rank's shard on its GPU: thin wrapper over sla_part_* (include/sla.h)."""

    def __init__(self, solver):
        self.solver = solver            # a KhoslaSolver whose CSR (the local rows) is uploaded / generated on device
        self.lib = _lib.load()
        self.ctx = solver._sync_device()
        self.device = torch.device("cuda", solver.device)
        self.stream = torch.cuda.ExternalStream(solver._context_stream(), device=self.device)
        self._words = self._cand = None

    @property
    def num_local_rows(self) -> int:
        return self.solver.num_rows()

    @property
    def num_cols(self) -> int:
        return self.solver.num_cols()

    def local_value_range(self):
        lo, hi, first = C.c_double(), C.c_double(), C.c_double()
        _lib.check(self.ctx, self.lib.sla_part_local_value_range(self.ctx, C.byref(lo), C.byref(hi), C.byref(first)))
        return lo.value, hi.value, first.value

    def begin(self, maximize, row_begin, global_rows, eps, gmin, gmax, gfirst):
        nan = float("nan")
        _lib.check(self.ctx, self.lib.sla_part_begin(self.ctx, _lib.ALGO_KHOSLA, int(bool(maximize)), row_begin, global_rows,
                                                     nan if eps is None else float(eps), gmin, gmax, gfirst))
        words, cand, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
        _lib.check(self.ctx, self.lib.sla_part_buffers(self.ctx, C.byref(words), C.byref(cand), C.byref(n)))
        self._words = torch.as_tensor(_DevBuf(words.value, n.value, "<i8"), device=self.device)
        self._cand = torch.as_tensor(_DevBuf(cand.value, n.value, "<f8"), device=self.device)

    def bid(self):
        _lib.check(self.ctx, self.lib.sla_part_bid(self.ctx))

    def claim(self):
        _lib.check(self.ctx, self.lib.sla_part_claim(self.ctx))

    def assign(self):
        q, d = C.c_uint32(), C.c_uint32()
        _lib.check(self.ctx, self.lib.sla_part_assign(self.ctx, C.byref(q), C.byref(d)))
        return q.value, d.value

    def words(self) -> torch.Tensor:
        return self._words

    def candidates(self) -> torch.Tensor:
        return self._cand

    # ---- sparse exchange ----
    def sparse_setup(self, world: int):
        send, recv, counts, cap = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        _lib.check(self.ctx, self.lib.sla_part_sparse_buffers(self.ctx, world, C.byref(send), C.byref(recv), C.byref(counts),
                                                              C.byref(cap)))
        self._send = torch.as_tensor(_DevBuf(send.value, 3 * cap.value, "<i8"), device=self.device)
        self._recv = torch.as_tensor(_DevBuf(recv.value, 3 * cap.value * world, "<i8"), device=self.device)
        self._counts = torch.as_tensor(_DevBuf(counts.value, world, "<i8"), device=self.device)

    def collect(self) -> int:
        n = C.c_uint32()
        _lib.check(self.ctx, self.lib.sla_part_collect(self.ctx, C.byref(n)))
        return n.value

    def send_list(self) -> torch.Tensor:
        return self._send

    def recv_lists(self) -> torch.Tensor:
        return self._recv

    def counts(self) -> torch.Tensor:
        return self._counts

    def apply_sparse(self, world: int, max_count: int):
        q, d = C.c_uint32(), C.c_uint32()
        _lib.check(self.ctx, self.lib.sla_part_apply_sparse(self.ctx, world, max_count, C.byref(q), C.byref(d)))
        return q.value, d.value

    def finish(self, download: bool = True):
        """Ends the solve; with download=False the local person_to_object slice and the replicated object_to_person /
        prices stay resident in HBM (fetch them later with solver.download_solution)."""
        n, m = self.num_local_rows, self.num_cols
        p2o = o2p = prices = None
        if download:
            from .solver import host_array
            p2o, o2p, prices = host_array(n, np.uint32), host_array(m, np.uint32), host_array(m, np.float64)
        st = SlaStats()
        _lib.check(self.ctx, self.lib.sla_part_finish(self.ctx, p2o.ctypes.data if download else None,
                                                      o2p.ctypes.data if download else None,
                                                      prices.ctypes.data if download else None, C.byref(st)))
        return p2o, o2p, prices, st.as_dict()

    def scalar_device(self):
        return self.device


class PartitionedKhoslaSolver:
    """Drives one row-partitioned solve; `engine` is this rank's shard (CudaShardEngine or a test model)."""

    def __init__(self, engine, group: Optional[dist.ProcessGroup] = None, exchange: str = "dense"):
        assert exchange in ("dense", "sparse")
        self.engine = engine
        self.group = group
        self.exchange = exchange
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.rounds = 0

    def _all_reduce(self, t, op):
        if self.world > 1:
            dist.all_reduce(t, op=op, group=self.group)
        return t

    def solve(self, maximize: bool = False, eps: Optional[float] = None, max_rounds: int = 1 << 30,
              download: bool = True) -> dict:
        eng = self.engine
        dev = eng.scalar_device()
        stream_ctx = torch.cuda.stream(eng.stream) if getattr(eng, "stream", None) is not None else _NullCtx()
        with stream_ctx:
            # ---- global shape: row offsets by an exclusive prefix sum of the shard sizes ----
            counts = torch.zeros(self.world, dtype=torch.int64, device=dev)
            counts[self.rank] = eng.num_local_rows
            self._all_reduce(counts, dist.ReduceOp.SUM)
            row_begin = int(counts[: self.rank].sum().item())
            global_rows = int(counts.sum().item())
            # ---- global value range and the sign-deciding first value (rank owning row 0) ----
            lo, hi, first = eng.local_value_range()
            rng = torch.tensor([-lo, hi], dtype=torch.float64, device=dev)
            self._all_reduce(rng, dist.ReduceOp.MAX)
            f = torch.tensor([first if row_begin == 0 and eng.num_local_rows > 0 else 0.0], dtype=torch.float64, device=dev)
            self._all_reduce(f, dist.ReduceOp.SUM)
            eng.begin(maximize, row_begin, global_rows, eps, -float(rng[0].item()), float(rng[1].item()), float(f[0].item()))

            self.rounds = 0
            qlen = torch.zeros(1, dtype=torch.int64, device=dev)
            if self.exchange == "sparse":
                eng.sparse_setup(self.world)
            while self.rounds < max_rounds:
                eng.bid()
                if self.exchange == "dense":
                    self._all_reduce(eng.words(), dist.ReduceOp.MAX)
                    eng.claim()
                    self._all_reduce(eng.candidates(), dist.ReduceOp.MAX)
                    local_q, _ = eng.assign()
                else:
                    local_q = self._sparse_round(eng, dev)
                self.rounds += 1
                qlen[0] = local_q
                self._all_reduce(qlen, dist.ReduceOp.SUM)
                if int(qlen.item()) == 0:
                    break
            p2o, o2p, prices, st = eng.finish(download) if download is False else eng.finish()
            tot = torch.tensor([st["num_unassigned"], st["bids"], st["bid_arcs"]], dtype=torch.int64, device=dev)
            self._all_reduce(tot, dist.ReduceOp.SUM)
            totals = [int(x) for x in tot.tolist()]      # read back on the engine's stream, behind the all-reduce
        st = dict(st)
        st["global_num_unassigned"], st["global_bids"], st["global_bid_arcs"] = totals
        st["row_begin"], st["global_rows"], st["rounds"] = row_begin, global_rows, self.rounds
        return dict(p2o=p2o, o2p=o2p, prices=prices, stats=st)


def _sparse_round(self, eng, dev):
    """One round of the sparse exchange: all-gather the per-rank winner counts, then the padded winner lists."""
    cnt = eng.collect()
    counts = eng.counts()
    mine = torch.tensor([cnt], dtype=torch.int64, device=dev)
    if self.world > 1:
        dist.all_gather(list(counts.view(self.world, 1).unbind(0)), mine, group=self.group)
    else:
        counts.copy_(mine)
    maxc = int(counts.max().item())
    if maxc > 0:
        send, recv = eng.send_list()[: 3 * maxc], eng.recv_lists()[: 3 * maxc * self.world]
        if self.world > 1:
            if dist.get_backend(self.group) == "nccl":
                dist.all_gather_into_tensor(recv, send, group=self.group)
            else:
                dist.all_gather(list(recv.view(self.world, 3 * maxc).unbind(0)), send, group=self.group)
        else:
            recv.copy_(send)
    local_q, _ = eng.apply_sparse(self.world, maxc)
    return local_q


PartitionedKhoslaSolver._sparse_round = _sparse_round


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
