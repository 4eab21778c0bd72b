"""GPU: the row-partitioned Khosla engine (sla_part_*).  One GPU is enough: world size 1 through the real driver, and
two / three shards stepped in lockstep inside one process with the MAX all-reduces done by hand (the profiling guide
forbids co-resident ranks that wait on each other on one GPU)."""
import numpy as np
import pytest
import torch

from helpers import random_sparse_instance

pytestmark = pytest.mark.gpu


def make_engine(sla, n_local, m, rp, c, v):
    from sparse_linear_assignment_b200.distributed import CudaShardEngine
    solver, _ = sla.KhoslaSolver.new(n_local, m, len(c))
    solver.load_csr(n_local, m, rp, c, v)
    return CudaShardEngine(solver)


@pytest.mark.parametrize("exchange", ["dense", "sparse"])
def test_world_size_one_equals_model(sla, oracle, exchange):
    from sparse_linear_assignment_b200.distributed import PartitionedKhoslaSolver
    rng = np.random.default_rng(4)
    n, m, k = 5000, 9000, 16
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=500)
    eng = make_engine(sla, n, m, rp, c, v)
    res = PartitionedKhoslaSolver(eng, exchange=exchange).solve(maximize=False, eps=None)
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v)
    assert np.array_equal(res["p2o"], ref["p2o"]) and np.array_equal(res["o2p"], ref["o2p"])
    assert np.array_equal(res["prices"], ref["prices"])
    assert res["stats"]["rounds"] == ref["stats"]["rounds"] and res["stats"]["global_bid_arcs"] == ref["stats"]["bid_arcs"]


@pytest.mark.parametrize("world,maximize,k", [(2, False, 16), (3, True, 7)])
def test_lockstep_shards_equal_model(sla, oracle, world, maximize, k):
    from sparse_linear_assignment_b200.distributed import shard_rows
    rng = np.random.default_rng(9 + world)
    n, m = 4001, 6000
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=500)
    engines, begins = [], []
    for r in range(world):
        b, cnt = shard_rows(n, world, r)
        a0, a1 = int(rp[b]), int(rp[b + cnt])
        engines.append(make_engine(sla, cnt, m, (rp[b:b + cnt + 1].astype(np.int64) - a0).astype(np.uint32), c[a0:a1], v[a0:a1]))
        begins.append(b)
    ranges = [e.local_value_range() for e in engines]
    gmin, gmax, gfirst = min(r[0] for r in ranges), max(r[1] for r in ranges), ranges[0][2]
    eps = 1.0 / (m + 1)
    for e, b in zip(engines, begins):
        e.begin(maximize, b, n, eps, gmin, gmax, gfirst)

    def all_reduce_max(tensors):
        torch.cuda.synchronize()
        red = tensors[0].clone()
        for t in tensors[1:]:
            red = torch.maximum(red, t)
        for t in tensors:
            t.copy_(red)
        torch.cuda.synchronize()

    rounds = 0
    while True:
        for e in engines:
            e.bid()
        all_reduce_max([e.words() for e in engines])
        for e in engines:
            e.claim()
        all_reduce_max([e.candidates() for e in engines])
        total = sum(e.assign()[0] for e in engines)
        rounds += 1
        if total == 0:
            break
        assert rounds < 100000
    outs = [e.finish() for e in engines]
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, maximize=maximize, eps=eps)
    assert np.array_equal(np.concatenate([o[0] for o in outs]), ref["p2o"])
    for o in outs:
        assert np.array_equal(o[1], ref["o2p"]) and np.array_equal(o[2], ref["prices"])
    assert rounds == ref["stats"]["rounds"]
    assert sum(o[3]["bid_arcs"] for o in outs) == ref["stats"]["bid_arcs"]
    assert sum(o[3]["num_unassigned"] for o in outs) == ref["stats"]["num_unassigned"]


@pytest.mark.parametrize("world,maximize,k", [(2, False, 16), (3, True, 5)])
def test_lockstep_shards_sparse_exchange_equal_model(sla, oracle, world, maximize, k):
    """Same as above with the sparse exchange: the all-gather of the winner lists is done by hand."""
    from sparse_linear_assignment_b200.distributed import shard_rows
    rng = np.random.default_rng(19 + world)
    n, m = 3001, 5000
    rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=500)
    engines, begins = [], []
    for r in range(world):
        b, cnt = shard_rows(n, world, r)
        a0, a1 = int(rp[b]), int(rp[b + cnt])
        engines.append(make_engine(sla, cnt, m, (rp[b:b + cnt + 1].astype(np.int64) - a0).astype(np.uint32), c[a0:a1], v[a0:a1]))
        begins.append(b)
    ranges = [e.local_value_range() for e in engines]
    gmin, gmax, gfirst = min(r[0] for r in ranges), max(r[1] for r in ranges), ranges[0][2]
    eps = 1.0 / (m + 1)
    for e, b in zip(engines, begins):
        e.begin(maximize, b, n, eps, gmin, gmax, gfirst)
        e.sparse_setup(world)
    rounds = 0
    while True:
        for e in engines:
            e.bid()
        counts = [e.collect() for e in engines]
        maxc = max(counts)
        torch.cuda.synchronize()
        for e in engines:
            e.counts().copy_(torch.tensor(counts, dtype=torch.int64))
            for r, src in enumerate(engines):
                e.recv_lists()[3 * maxc * r: 3 * maxc * (r + 1)].copy_(src.send_list()[: 3 * maxc])
        torch.cuda.synchronize()
        total = sum(e.apply_sparse(world, maxc)[0] for e in engines)
        rounds += 1
        if total == 0:
            break
        assert rounds < 100000
    outs = [e.finish() for e in engines]
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, maximize=maximize, eps=eps)
    assert np.array_equal(np.concatenate([o[0] for o in outs]), ref["p2o"])
    for o in outs:
        assert np.array_equal(o[1], ref["o2p"]) and np.array_equal(o[2], ref["prices"])
    assert rounds == ref["stats"]["rounds"]
    assert sum(o[3]["bid_arcs"] for o in outs) == ref["stats"]["bid_arcs"]
