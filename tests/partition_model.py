"""Test infrastructure: a pure-numpy model of one rank's shard of the row-partitioned Khosla engine, with the same
interface as sparse_linear_assignment_b200.distributed.CudaShardEngine, so that the host-side driver (collectives,
termination, offsets) can be exercised on CPU with the gloo backend."""
import numpy as np
import torch

NONE = 0xFFFFFFFF


class ModelShardEngine:
    stream = None

    def __init__(self, oracle, num_cols, row_ptr, cols, vals):
        self.O = oracle
        self.m = int(num_cols)
        self.rp = np.asarray(row_ptr, dtype=np.int64)
        self.cols = np.asarray(cols, dtype=np.int64)
        self.vals = np.asarray(vals, dtype=np.float64)
        self.n = self.rp.size - 1

    num_local_rows = property(lambda self: self.n)
    num_cols = property(lambda self: self.m)

    def scalar_device(self):
        return torch.device("cpu")

    def local_value_range(self):
        return float(self.vals.min()), float(self.vals.max()), float(self.vals[0])

    def begin(self, maximize, row_begin, global_rows, eps, gmin, gmax, gfirst):
        self.row_begin, self.global_rows = row_begin, global_rows
        flip = bool(maximize) != (gfirst >= 0.0)
        self.sign = -1.0 if flip else 1.0
        wmin, wmax = (-gmax, -gmin) if flip else (gmin, gmax)
        self.eps = 1.0 / self.m if eps is None else eps
        self.threshold = (self.m / 2.0) * (wmax - wmin + self.eps)
        self.pbits = max(int(global_rows - 1 if global_rows > 1 else 1).bit_length(), 1)
        self.prices = np.zeros(self.m)
        self.o2p = np.full(self.m, NONE, dtype=np.uint32)
        self.p2o = np.full(self.n, NONE, dtype=np.uint32)
        self._words = torch.zeros(self.m, dtype=torch.int64)
        self._cand = torch.full((self.m,), float("-inf"), dtype=torch.float64)
        self.queue = list(range(self.n))
        self.next_queue = []
        self.slots = []
        self.dropped = self.bids = self.arcs = self.rounds = 0
        self.flip = flip

    def words(self):
        return self._words

    def candidates(self):
        return self._cand

    def bid(self):
        self.slots = []
        w = self._words.numpy()
        for i in self.queue:
            a, b = self.rp[i], self.rp[i + 1]
            best = second = value = float("-inf")
            jbest = 0
            for g in range(a, b):
                v = self.sign * self.vals[g] if self.sign < 0 else self.vals[g]
                profit = v - self.prices[self.cols[g]]
                if profit > best:
                    jbest, second, best, value = int(self.cols[g]), best, profit, v
                elif profit > second:
                    second = profit
            self.arcs += int(b - a)
            self.bids += 1
            if self.prices[jbest] > self.threshold:
                self.dropped += 1
                self.slots.append((i, NONE, 0.0))
                continue
            bid = value - second + self.eps if np.isfinite(second) else self.prices[jbest] + self.eps
            self.slots.append((i, jbest, bid))
            if bid == bid:
                word = self.O.pack_bid(bid, i + self.row_begin, self.pbits)
                if word > w[jbest]:
                    w[jbest] = word

    def claim(self):
        w, c = self._words.numpy(), self._cand.numpy()
        self.next_queue = []
        for i, j, bid in self.slots:
            if j == NONE:
                continue
            if bid == bid and int(w[j]) == self.O.pack_bid(bid, i + self.row_begin, self.pbits):
                c[j] = bid
            else:
                self.next_queue.append(i)

    def assign(self):
        w, c = self._words.numpy(), self._cand.numpy()
        pmask = (1 << self.pbits) - 1
        for j in np.nonzero(w)[0]:
            person = pmask - (int(w[j]) & pmask)
            prev = int(self.o2p[j])
            self.prices[j] = c[j]
            self.o2p[j] = person
            w[j] = 0
            c[j] = float("-inf")
            if 0 <= person - self.row_begin < self.n:
                self.p2o[person - self.row_begin] = j
            if prev != NONE and 0 <= prev - self.row_begin < self.n:
                self.p2o[prev - self.row_begin] = NONE
                self.next_queue.append(prev - self.row_begin)
        self.queue = self.next_queue
        self.rounds += 1
        return len(self.queue), self.dropped

    # ---- sparse exchange (same interface as CudaShardEngine) ----
    def sparse_setup(self, world):
        cap = (self.global_rows + world - 1) // world + 1
        self._send = torch.zeros(3 * cap, dtype=torch.int64)
        self._recv = torch.zeros(3 * cap * world, dtype=torch.int64)
        self._counts_t = torch.zeros(world, dtype=torch.int64)

    def send_list(self):
        return self._send

    def recv_lists(self):
        return self._recv

    def counts(self):
        return self._counts_t

    def collect(self):
        w = self._words.numpy()
        send = self._send.numpy()
        self.next_queue = []
        n = 0
        for i, j, bid in self.slots:
            if j == NONE:
                continue
            word = self.O.pack_bid(bid, i + self.row_begin, self.pbits) if bid == bid else 0
            if bid == bid and int(w[j]) == word:
                send[3 * n:3 * n + 3] = (j, word, np.float64(bid).view(np.int64))
                n += 1
            else:
                self.next_queue.append(i)
        for i, j, bid in self.slots:
            if j != NONE:
                w[j] = 0
        return n

    def apply_sparse(self, world, maxc):
        w = self._words.numpy()
        recv = self._recv.numpy()
        cnt = self._counts_t.numpy()
        pmask = (1 << self.pbits) - 1
        entries = [tuple(int(x) for x in recv[3 * (r * maxc + e):3 * (r * maxc + e) + 3]) for r in range(world) for e in range(int(cnt[r]))]
        for j, word, _ in entries:
            if word > w[j]:
                w[j] = word
        for j, word, bits in entries:
            person = pmask - (word & pmask)
            if int(w[j]) == word:
                prev = int(self.o2p[j])
                self.prices[j] = np.int64(bits).view(np.float64)
                self.o2p[j] = person
                if 0 <= person - self.row_begin < self.n:
                    self.p2o[person - self.row_begin] = j
                if prev != NONE and 0 <= prev - self.row_begin < self.n:
                    self.p2o[prev - self.row_begin] = NONE
                    self.next_queue.append(prev - self.row_begin)
            elif 0 <= person - self.row_begin < self.n:
                self.next_queue.append(person - self.row_begin)
        for j, _, _ in entries:
            w[j] = 0
        self.queue = self.next_queue
        self.rounds += 1
        return len(self.queue), self.dropped

    def finish(self, download=True):
        st = dict(num_unassigned=self.dropped, nits=self.bids, bids=self.bids, bid_arcs=self.arcs, rounds=self.rounds,
                  dropped=self.dropped, eps=self.eps, values_negated=int(self.flip))
        return self.p2o, self.o2p, self.prices, st
