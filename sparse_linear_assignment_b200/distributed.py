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

__all__ = ["PartitionedKhoslaSolver", "CudaShardEngine", "shard_rows", "MeshKhoslaSolver", "MeshShard", "mesh_lockstep_solve",
           "object_shard"]


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


# =====================================================================================================================
# Mesh engine: owner-partitioned objects, bids pushed into the owner's HBM over NVLink peer mappings (csrc/sla_mesh.cuh)
# =====================================================================================================================
def object_shard(num_cols: int, world: int) -> int:
    """Objects per rank of the mesh engine: the power of two at or above ceil(num_cols / world)."""
    per = (num_cols + world - 1) // world
    s = 1
    while s < per:
        s <<= 1
    return s


class MeshShard:
    """This rank's shard in the mesh engine: thin wrapper over sla_mesh_* (include/sla.h).  `solver` is a KhoslaSolver
    whose CSR holds the local rows (uploaded or generated with the GLOBAL number of columns)."""

    def __init__(self, solver, rank: int, world: int, row_begins):
        self.solver, self.rank, self.world = solver, int(rank), int(world)
        self.lib = _lib.load()
        self.ctx = solver._sync_device()
        self.row_begins = np.ascontiguousarray(row_begins, dtype=np.uint32)
        assert self.row_begins.size == world + 1
        block, nbytes = C.c_void_p(), C.c_size_t()
        _lib.check(self.ctx, self.lib.sla_mesh_create(self.ctx, self.rank, self.world, self.row_begins.ctypes.data,
                                                      solver.num_cols(), C.byref(block), C.byref(nbytes)))
        self.block, self.block_bytes = block.value, nbytes.value
        self._imported = []

    def export_handle(self) -> bytes:
        h = (C.c_ubyte * 64)()
        rc = self.lib.sla_ipc_export(self.block, h)
        if rc != _lib.SLA_OK:
            raise _lib.SlaError(rc, "cudaIpcGetMemHandle failed for the mesh block")
        return bytes(h)

    def connect_handles(self, handles):
        """handles[g]: rank g's 64-byte IPC handle (the entry of this rank is ignored)."""
        ptrs = (C.c_void_p * self.world)()
        for g in range(self.world):
            if g == self.rank:
                ptrs[g] = self.block
                continue
            buf = (C.c_ubyte * 64).from_buffer_copy(handles[g])
            out = C.c_void_p()
            rc = self.lib.sla_ipc_import(self.solver.device, buf, C.byref(out))
            if rc != _lib.SLA_OK:
                msg = self.lib.sla_last_error(None)
                raise _lib.SlaError(rc, f"cannot map rank {g}'s mesh block: " + (msg.decode() if msg else ""))
            self._imported.append(out.value)
            ptrs[g] = out.value
        _lib.check(self.ctx, self.lib.sla_mesh_connect(self.ctx, ptrs, None))

    def connect_pointers(self, blocks, devices=None):
        """Same process: the other shards' block pointers (and their devices, when they differ)."""
        ptrs = (C.c_void_p * self.world)(*blocks)
        devs = (C.c_int * self.world)(*devices) if devices is not None else None
        _lib.check(self.ctx, self.lib.sla_mesh_connect(self.ctx, ptrs, devs))

    def local_value_range(self):
        lo, hi, first = C.c_double(), C.c_double(), C.c_double()
        _lib.check(self.ctx, self.lib.sla_part_local_value_range(self.ctx, C.byref(lo), C.byref(hi), C.byref(first)))
        return lo.value, hi.value, first.value

    def begin(self, maximize, eps, gmin, gmax, gfirst):
        _lib.check(self.ctx, self.lib.sla_mesh_begin(self.ctx, int(bool(maximize)), float("nan") if eps is None else float(eps),
                                                     gmin, gmax, gfirst))

    def solve(self):
        _lib.check(self.ctx, self.lib.sla_mesh_solve(self.ctx))

    def phase(self, which: int):
        _lib.check(self.ctx, self.lib.sla_mesh_phase(self.ctx, which))

    def poll(self):
        d, r, q = C.c_int(), C.c_uint32(), C.c_uint32()
        _lib.check(self.ctx, self.lib.sla_mesh_poll(self.ctx, C.byref(d), C.byref(r), C.byref(q)))
        return bool(d.value), r.value, q.value

    def owned(self):
        s, n, f, r = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        _lib.check(self.ctx, self.lib.sla_mesh_owned(self.ctx, C.byref(s), C.byref(n), C.byref(f), C.byref(r)))
        return dict(shard_objects=s.value, num_owned=n.value, first_object=f.value, first_row=r.value)

    def finish(self, download: bool = True):
        """Returns (person_to_object of the local rows, object_to_person and prices of the owned objects, stats)."""
        own = self.owned()
        n = self.solver.num_rows()
        p2o = o2p = prices = None
        if download:
            # page-locked result buffers are kept across solves (cudaMallocHost costs milliseconds): the returned arrays
            # are overwritten by the next finish(download=True) of this shard
            from .solver import host_array
            if getattr(self, "_out_shape", None) != (n, own["num_owned"]):
                self._out = (host_array(n, np.uint32), host_array(max(own["num_owned"], 1), np.uint32)[: own["num_owned"]],
                             host_array(max(own["num_owned"], 1), np.float64)[: own["num_owned"]])
                self._out_shape = (n, own["num_owned"])
            p2o, o2p, prices = self._out
        st = SlaStats()
        _lib.check(self.ctx, self.lib.sla_mesh_finish(self.ctx, p2o.ctypes.data if download else None,
                                                      o2p.ctypes.data if download and own["num_owned"] else None,
                                                      prices.ctypes.data if download and own["num_owned"] else None,
                                                      C.byref(st)))
        return p2o, o2p, prices, st.as_dict()

    def round1_ms(self) -> float:
        out = C.c_float()
        _lib.check(self.ctx, self.lib.sla_mesh_round1_ms(self.ctx, C.byref(out)))
        return float(out.value)

    def objective(self) -> float:
        out = C.c_double()
        _lib.check(self.ctx, self.lib.sla_mesh_objective(self.ctx, C.byref(out)))
        return float(out.value)

    def close(self):
        for ptr in self._imported:
            self.lib.sla_ipc_release(self.solver.device, ptr)
        self._imported = []


def _global_range(shards_or_ranges):
    lo = min(r[0] for r in shards_or_ranges)
    hi = max(r[1] for r in shards_or_ranges)
    return lo, hi


def mesh_lockstep_solve(shards, maximize=False, eps=None, max_rounds=1 << 20):
    """Drives all ranks of a mesh solve from ONE host thread, one kernel at a time (every rank's phase k before any rank's
    phase k + 1) -- the way to run G ranks on fewer than G GPUs, where kernels that wait for each other's flags must never
    be resident together (tests; same kernels, same peer pointers, same results as the concurrent path).
    Returns the assembled global solution."""
    world = len(shards)
    ranges = [s.local_value_range() for s in shards]
    gmin, gmax = _global_range(ranges)
    gfirst = ranges[0][2]
    for s in shards:
        s.begin(maximize, eps, gmin, gmax, gfirst)
    rounds = 0
    while rounds < max_rounds:
        for which in range(4):
            for s in shards:
                s.phase(which)
        states = [s.poll() for s in shards]
        assert len({st[0] for st in states}) == 1, "ranks disagree about the end of the solve"
        if states[0][0]:
            break
        rounds += 1
    return mesh_assemble([s.finish() for s in shards], [s.owned() for s in shards], shards[0].solver.num_cols())


def mesh_assemble(parts, owned, num_cols):
    """Global vectors from the ranks' slices: parts[g] = (p2o_local, o2p_owned, prices_owned, stats)."""
    p2o = np.concatenate([p[0] for p in parts])
    o2p = np.full(num_cols, 0xFFFFFFFF, dtype=np.uint32)
    prices = np.zeros(num_cols, dtype=np.float64)
    for (_, o, pr, _), own in zip(parts, owned):
        a, n = own["first_object"], own["num_owned"]
        o2p[a:a + n] = o
        prices[a:a + n] = pr
    stats = dict(parts[0][3])
    for key in ("num_unassigned", "bids", "bid_arcs", "dropped", "nits"):
        stats[key] = int(sum(p[3][key] for p in parts))
    return dict(p2o=p2o, o2p=o2p, prices=prices, stats=stats)


class MeshKhoslaSolver:
    """One KhoslaSolver instance over the ranks of a process group, one process per GPU (KhoslaSolver::solve,
    ksparse.rs:153-251, for an instance whose rows are spread over the GPUs of one NVLink domain).  torch.distributed is
    used for the plumbing only: shard sizes, the exchange of the 64-byte IPC handles, the global value range and the final
    totals.  The rounds themselves never touch the host or a collective (csrc/sla_mesh.cuh)."""

    def __init__(self, solver, group: Optional[dist.ProcessGroup] = None, shard_factory=None):
        self.solver = solver
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        backend = dist.get_backend(group) if dist.is_initialized() else "none"
        self.dev = torch.device("cuda", solver.device) if backend == "nccl" else torch.device("cpu")
        self.shard = None
        self.row_begins = None
        self.shard_factory = shard_factory or MeshShard     # the CPU tests put a numpy model of the protocol here

    def _gather_i64(self, value: int):
        t = torch.zeros(self.world, dtype=torch.int64, device=self.dev)
        t[self.rank] = int(value)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return [int(x) for x in t.tolist()]

    def setup(self):
        """Shard layout, this rank's mesh block, peer mappings of every other rank's block.  A failure on any rank (no
        peer access, out of memory, ...) is agreed upon collectively and raised on EVERY rank, so that callers can fall
        back together instead of leaving the others in a collective."""
        counts = self._gather_i64(self.solver.num_rows())
        self.row_begins = np.concatenate([[0], np.cumsum(counts)]).astype(np.uint32)
        problem = ""
        mine = torch.zeros(64, dtype=torch.uint8)
        try:
            self.shard = self.shard_factory(self.solver, self.rank, self.world, self.row_begins)
            mine = torch.frombuffer(bytearray(self.shard.export_handle()), dtype=torch.uint8)
        except Exception as e:          # noqa: BLE001 -- reported through the collective below
            problem = f"rank {self.rank}: {e}"
        allh = torch.zeros(self.world * 64, dtype=torch.uint8, device=self.dev)
        allh[self.rank * 64:(self.rank + 1) * 64] = mine.to(self.dev)
        if self.world > 1:
            dist.all_reduce(allh, op=dist.ReduceOp.SUM, group=self.group)      # disjoint slices: a sum is a gather
        if not problem:
            try:
                raw = bytes(allh.cpu().numpy().tobytes())
                self.shard.connect_handles([raw[g * 64:(g + 1) * 64] for g in range(self.world)])
            except Exception as e:      # noqa: BLE001
                problem = f"rank {self.rank}: {e}"
        ok = torch.tensor([0 if problem else 1], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            # (also the barrier: every rank has mapped every block before anybody starts a solve)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        if int(ok.item()) == 0:
            self.shard = None
            raise _lib.SlaError(_lib.SLA_ERR_CUDA, "mesh setup failed on at least one rank" + (f" ({problem})" if problem else ""))
        return self

    def _global_range(self):
        """Global (min, max, first value) of the instance: one small all-reduce, cached while the shards stay resident."""
        lo, hi, first = self.shard.local_value_range()
        rng = torch.tensor([-lo, hi, first if self.rank == 0 else 0.0], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            mx = rng[:2].clone()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
            f = rng[2:].clone()
            dist.all_reduce(f, op=dist.ReduceOp.SUM, group=self.group)
            rng = torch.cat([mx, f])
        return -float(rng[0].item()), float(rng[1].item()), float(rng[2].item())

    def solve(self, maximize: bool = False, eps: Optional[float] = None, download: bool = True, gather: bool = False,
              totals: bool = True) -> dict:
        if self.shard is None:
            self.setup()
        sh = self.shard
        # a KhoslaSolver with host storage: mirror a changed CSR into HBM, with the in-place sign normalisation of the
        # host copy of `values` that init_solve performs (solver.rs:209-216); every rank decides by its own first value
        # (inputs with mixed signs are not supported by the reference either)
        sync = getattr(self.solver, "_sync_device", None)
        resynced = False
        if sync is not None and getattr(self.solver, "_dirty", False) and not getattr(self.solver, "_device_only", False):
            sync(maximize)
            self.solver._pre_negated = False
            resynced = True
        if resynced or getattr(self, "_range", None) is None:
            self._range = self._global_range()
        gmin, gmax, gfirst = self._range
        sh.begin(maximize, eps, gmin, gmax, gfirst)
        sh.solve()                                   # graphs of rounds; the ranks meet only in the kernels' flag barriers
        p2o, o2p, prices, st = sh.finish(download)
        st = dict(st)
        if totals:
            tot = torch.tensor([st["num_unassigned"], st["bids"], st["bid_arcs"]], dtype=torch.int64, device=self.dev)
            if self.world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
            st["global_num_unassigned"], st["global_bids"], st["global_bid_arcs"] = [int(x) for x in tot.tolist()]
        own = sh.owned()
        out = dict(p2o=p2o, o2p=o2p, prices=prices, stats=st, owned=own, row_begin=int(self.row_begins[self.rank]),
                   global_rows=int(self.row_begins[-1]))
        if gather and download:
            out.update(self._gather_solution(p2o, o2p, prices, own))
        return out

    def objective(self) -> float:
        """get_objective (solver.rs:110-142) of the last solve: the ranks' shares added up (exact for integer weights)."""
        t = torch.tensor([self.shard.objective()], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return float(t.item())

    def _gather_solution(self, p2o, o2p, prices, own):
        """All ranks' slices on every rank (tests and small instances; large ones keep their slices)."""
        n, m = int(self.row_begins[-1]), self.solver.num_cols()
        gp = torch.zeros(n, dtype=torch.int64, device=self.dev)
        a = int(self.row_begins[self.rank])
        gp[a:a + p2o.size] = torch.from_numpy(p2o.astype(np.int64)).to(self.dev) + 1      # 0 = not mine
        go = torch.zeros(m, dtype=torch.int64, device=self.dev)
        gpr = torch.zeros(m, dtype=torch.float64, device=self.dev)
        f, k = own["first_object"], own["num_owned"]
        go[f:f + k] = torch.from_numpy(o2p.astype(np.int64)).to(self.dev) + 1
        gpr[f:f + k] = torch.from_numpy(np.ascontiguousarray(prices)).to(self.dev)
        if self.world > 1:
            for t in (gp, go, gpr):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return dict(global_p2o=(gp - 1).cpu().numpy().astype(np.uint32), global_o2p=(go - 1).cpu().numpy().astype(np.uint32),
                    global_prices=gpr.cpu().numpy())

    def close(self):
        if self.shard is not None:
            self.shard.close()
