"""Test infrastructure: a numpy model of ONE rank of the mesh engine (csrc/sla_mesh.cuh) with the interface of
sparse_linear_assignment_b200.distributed.MeshShard, so that MeshKhoslaSolver's host side (shard layout, exchange of the
block handles, global value range, totals, gathering of the slices) and the protocol itself -- owner-partitioned objects,
bids pushed to the owner, reply bits, evictions pushed to the rank that holds the person, termination on the sum of the
queue lengths -- run on CPU over gloo.  Where the CUDA engine stores into a peer's memory, the model hands the same
records to the peer through the process group."""
import os

import numpy as np
import torch.distributed as dist

NONE = 0xFFFFFFFF


class _HostCsr:
    """Stands in for the KhoslaSolver the CUDA shard wraps: only what MeshKhoslaSolver touches."""
    device = 0

    def __init__(self, num_cols, row_ptr, cols, vals):
        self.m = int(num_cols)
        self.rp = np.asarray(row_ptr, dtype=np.int64)
        self.cols = np.asarray(cols, dtype=np.int64)
        self.vals = np.asarray(vals, dtype=np.float64)

    def num_rows(self):
        return self.rp.size - 1

    def num_cols(self):
        return self.m


class ModelMeshShard:
    def __init__(self, solver, rank, world, row_begins, oracle=None, group=None):
        from sparse_linear_assignment_b200.distributed import object_shard
        self.solver, self.rank, self.world = solver, rank, world
        self.row_begins = np.asarray(row_begins, dtype=np.int64)
        self.O, self.group = oracle, group
        self.n, self.m = solver.num_rows(), solver.num_cols()
        self.shard = object_shard(self.m, world)
        self.first_obj = min(rank * self.shard, self.m)
        self.owned_n = max(0, min(self.shard, self.m - rank * self.shard))
        self.connected = False

    # ---- handles: 64 bytes that identify the exporting rank (the CUDA shard's are cudaIpcMemHandle_t) ----
    def export_handle(self):
        return (b"mesh-model" + bytes([self.rank]) + os.getpid().to_bytes(4, "little")).ljust(64, b"\0")

    def connect_handles(self, handles):
        assert len(handles) == self.world and all(len(h) == 64 for h in handles)
        assert [h[10] for h in handles] == list(range(self.world)), "handles arrived out of rank order"
        assert handles[self.rank] == self.export_handle()
        self.connected = True

    def local_value_range(self):
        v = self.solver.vals
        return float(v.min()), float(v.max()), float(v[0])

    def owned(self):
        return dict(shard_objects=self.shard, num_owned=self.owned_n, first_object=self.first_obj,
                    first_row=int(self.row_begins[self.rank]))

    def begin(self, maximize, eps, gmin, gmax, gfirst):
        assert self.connected
        flip = bool(maximize) != (gfirst >= 0.0)
        self.sign = -1.0 if flip else 1.0
        wmin, wmax = (-gmax, -gmin) if flip else (gmin, gmax)
        self.eps = 1.0 / self.m if eps is None else eps
        self.threshold = (self.m / 2.0) * (wmax - wmin + self.eps)
        total_rows = int(self.row_begins[-1])
        self.pbits = max(int(total_rows - 1 if total_rows > 1 else 1).bit_length(), 1)
        self.best = np.zeros(self.owned_n, dtype=np.uint64)
        self.price = np.zeros(self.owned_n)
        self.owner = np.full(self.owned_n, NONE, dtype=np.uint32)
        self.p2o = np.full(self.n, NONE, dtype=np.uint32)
        self.queue = list(range(self.n))
        self.dropped = self.bids = self.arcs = self.rounds = 0
        self.flip = flip

    def _exchange(self, outboxes):
        """outboxes[g]: records for rank g -> list over senders of the records addressed to this rank."""
        if self.world == 1:
            return [outboxes[0]]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, outboxes, group=self.group)
        return [gathered[src][self.rank] for src in range(self.world)]

    def _all_prices(self):
        if self.world == 1:
            return self.price
        parts = [None] * self.world
        dist.all_gather_object(parts, self.price, group=self.group)
        return np.concatenate(parts)

    def solve(self):
        rb = int(self.row_begins[self.rank])
        s = self.solver
        while True:
            prices = self._all_prices()                 # (the CUDA kernel gathers single prices from the owners instead)
            # ---- K1: bids, bucketed by owner ----
            out = [[] for _ in range(self.world)]
            slots = []
            for i in self.queue:
                a, b = s.rp[i], s.rp[i + 1]
                best = second = value = float("-inf")
                jbest = 0
                for g in range(a, b):
                    v = -s.vals[g] if self.sign < 0 else s.vals[g]
                    profit = v - prices[s.cols[g]]
                    if profit > best:
                        jbest, second, best, value = int(s.cols[g]), best, profit, v
                    elif profit > second:
                        second = profit
                self.arcs += int(b - a)
                self.bids += 1
                if prices[jbest] > self.threshold:
                    self.dropped += 1
                    slots.append((i, NONE, None))
                    continue
                bid = value - second + self.eps if np.isfinite(second) else prices[jbest] + self.eps
                g = jbest // self.shard
                slots.append((i, g, len(out[g])))
                out[g].append((jbest - g * self.shard, i + rb, bid))
            inbox = self._exchange(out)
            # ---- K2 / K3 at the owner: maximum word per object, winners, reply bits, evictions ----
            for entries in inbox:
                for jl, person, bid in entries:
                    w = self.O.pack_bid(bid, person, self.pbits)
                    if w > int(self.best[jl]):
                        self.best[jl] = w
            replies = [[] for _ in range(self.world)]
            evict = [[] for _ in range(self.world)]
            for src, entries in enumerate(inbox):
                for jl, person, bid in entries:
                    won = int(self.best[jl]) == self.O.pack_bid(bid, person, self.pbits)
                    replies[src].append(won)
                    if won:
                        prev = int(self.owner[jl])
                        self.price[jl], self.owner[jl], self.best[jl] = bid, person, 0
                        if prev != NONE:
                            evict[int(np.searchsorted(self.row_begins, prev, side="right") - 1)].append(prev)
            got_replies = self._exchange(replies)
            got_evicted = self._exchange(evict)
            # ---- K4 at the bidder: outcomes, intake of the evicted, next queue ----
            nxt = []
            for i, g, pos in slots:
                if g == NONE:
                    continue
                if got_replies[g][pos]:
                    self.p2o[i] = g * self.shard + out[g][pos][0]
                else:
                    nxt.append(i)
            for lst in got_evicted:
                for person in lst:
                    self.p2o[person - rb] = NONE
                    nxt.append(person - rb)
            self.queue = nxt
            self.rounds += 1
            totals = [None] * self.world
            if self.world > 1:
                dist.all_gather_object(totals, len(nxt), group=self.group)
            else:
                totals = [len(nxt)]
            if sum(totals) == 0:
                break

    def finish(self, download=True):
        st = dict(num_unassigned=self.dropped, nits=self.bids, bids=self.bids, bid_arcs=self.arcs, rounds=self.rounds,
                  dropped=self.dropped, eps=self.eps, values_negated=int(self.flip), ms_solve=0.0, graph_launches=0,
                  kernel_launches=0)
        return self.p2o, self.owner, self.price, st

    def close(self):
        pass
