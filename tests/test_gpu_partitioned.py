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


# ---- mesh engine: owner-partitioned objects, bids pushed into the owner's memory through peer pointers -----------------
def make_mesh_shards(sla, n, m, rp, c, v, world, options=None):
    from sparse_linear_assignment_b200.distributed import MeshShard, shard_rows
    begins = [shard_rows(n, world, r)[0] for r in range(world)] + [n]
    shards = []
    for r in range(world):
        b, cnt = begins[r], begins[r + 1] - begins[r]
        a0, a1 = int(rp[b]), int(rp[b + cnt])
        solver, _ = sla.KhoslaSolver.new(cnt, m, max(a1 - a0, 1))
        for k_, v_ in (options or {}).items():
            solver.set_option(k_, v_)
        solver.load_csr(cnt, m, (rp[b:b + cnt + 1].astype(np.int64) - a0).astype(np.uint32), c[a0:a1], v[a0:a1].copy())
        shards.append(MeshShard(solver, r, world, begins))
    for s in shards:
        s.connect_pointers([t.block for t in shards])
    return shards


def assert_mesh_equals_model(res, ref):
    assert np.array_equal(res["p2o"], ref["p2o"]), "person_to_object differs from the model"
    assert np.array_equal(res["o2p"], ref["o2p"]), "object_to_person differs from the model"
    assert np.array_equal(res["prices"], ref["prices"]), "prices differ from the model (bit-exact f64)"
    for key in ("num_unassigned", "bids", "bid_arcs", "dropped", "rounds"):
        assert res["stats"][key] == ref["stats"][key], (key, res["stats"][key], ref["stats"][key])


@pytest.mark.parametrize("world,n,m,k,maximize,ragged", [
    (2, 4001, 6000, 16, False, False),       # uniform degree: the LDG.256 scan, two lanes per row
    (3, 3000, 7000, 8, True, False),         # one lane per row, objects split 4096 / 2904 / 0: a rank that owns nothing
    (4, 5003, 7000, 24, False, False),       # three of four lanes busy
    (8, 2500, 9000, 40, False, False),       # eight ranks, K > 32
    (3, 3001, 5000, 7, False, True),         # ragged CSR: the masked 128-bit scan
    (2, 2, 5, 3, False, True),               # one person per rank
])
def test_mesh_lockstep_equals_model(sla, oracle, world, n, m, k, maximize, ragged):
    """All ranks of the mesh engine on ONE GPU, stepped in lockstep (sla_mesh_phase): every rank's block is reached
    through the same peer-pointer tables as across GPUs.  Bit-identical to the sequential model: assignment, prices,
    rounds and work counters; solved twice to cover the reuse of the blocks and the running barrier epoch."""
    from sparse_linear_assignment_b200.distributed import mesh_lockstep_solve
    rng = np.random.default_rng(100 * world + k)
    if ragged:
        # rows of 1 .. k arcs around a planted perfect matching (without one the plain Khosla rounds run until the price
        # threshold: billions of rounds for persons that share a single object)
        counts = rng.integers(1, k + 1, size=n)
        rp = np.zeros(n + 1, dtype=np.uint32)
        rp[1:] = np.cumsum(counts)
        perm = rng.permutation(m)[:n]
        rows = []
        for i in range(n):
            others = rng.choice(m - 1, size=int(counts[i]) - 1, replace=False)
            others = others + (others >= perm[i])
            rows.append(np.sort(np.concatenate([[perm[i]], others])))
        c = np.concatenate(rows).astype(np.uint32)
        v = rng.integers(1, 500, size=int(rp[-1])).astype(np.float64)
    else:
        rp, c, v = random_sparse_instance(rng, n, m, k, integer=True, lo=1, hi=500)
    eps = 0.5 if ragged else 1.0 / (m + 1)   # (rows with a single arc: a small eps means price wars of a million rounds)
    shards = make_mesh_shards(sla, n, m, rp, c, v, world)
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, maximize=maximize, eps=eps, khosla_scaling=False)
    assert ref["stats"]["rounds"] < 5000
    for _ in range(2):
        res = mesh_lockstep_solve(shards, maximize=maximize, eps=eps)
        assert_mesh_equals_model(res, ref)


def test_mesh_lockstep_large_rounds_u16_and_f64(sla, oracle):
    """Rounds large enough for several staging chunks per block and the u16 value mirror (70,000 persons per rank), then
    the same instance with real-valued weights (f64 scan); against the one-GPU engine bit for bit."""
    from sparse_linear_assignment_b200.distributed import mesh_lockstep_solve
    n, m, k, world = 140_000, 400_000, 16, 2
    rp, c, v = sla.generators.kregular_host(n, m, k, seed=11)
    for vals in (v, v + 0.125):
        single, z = sla.KhoslaSolver.new(n, m, n * k)
        single.load_csr(n, m, rp, c, vals.copy())
        single.solve(z, False, None)
        shards = make_mesh_shards(sla, n, m, rp, c, vals, world)
        if vals is v:
            assert all(s.solver.scan_value_bytes() == 2 for s in shards)
        res = mesh_lockstep_solve(shards, maximize=False, eps=None)
        assert np.array_equal(res["p2o"], z.person_to_object) and np.array_equal(res["o2p"], z.object_to_person)
        assert np.array_equal(res["prices"], single.prices())
        for key in ("bids", "bid_arcs", "rounds", "num_unassigned"):
            assert res["stats"][key] == single.last_stats[key], key


def test_mesh_threshold_drops_match_model(sla, oracle):
    """An instance without a perfect matching: persons dropped at the price threshold (ksparse.rs:218-220) are counted
    per rank and add up to the model's num_unassigned."""
    from sparse_linear_assignment_b200.distributed import mesh_lockstep_solve
    rng = np.random.default_rng(5)
    n, m, k = 60, 70, 3
    rp = np.arange(0, n * k + 1, k, dtype=np.uint32)
    c = np.sort(rng.integers(0, 20, size=(n, k)), axis=1).astype(np.uint32).reshape(-1)      # 60 persons want 20 objects
    v = rng.integers(1, 50, size=n * k).astype(np.float64)
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, khosla_scaling=False)
    assert ref["stats"]["num_unassigned"] >= 40
    res = mesh_lockstep_solve(make_mesh_shards(sla, n, m, rp, c, v, 3))
    assert_mesh_equals_model(res, ref)


@pytest.mark.parametrize("shape", ["regular_u16", "regular_f64", "ragged", "k64"])
@pytest.mark.parametrize("tail", [1, 0])
def test_mesh_concurrent_path_world_one_equals_model(sla, oracle, shape, tail):
    """The production path of the mesh engine -- sla_mesh_solve: rounds captured into CUDA graphs, barriers waited for in
    the producing kernel's last block, the persistent one-block tail engine for the short rounds (option mesh_tail) -- with
    a world of ONE rank (one GPU is enough: the only flags it waits for are its own).  Bit-identical to the model; solved
    three times on the resident shard (the graph length learned from the previous solve, the running barrier epoch)."""
    from sparse_linear_assignment_b200.distributed import MeshShard, mesh_assemble
    rng = np.random.default_rng(77)
    if shape == "ragged":
        n, m = 30_000, 50_000
        counts = rng.integers(2, 12, size=n)
        rp = np.zeros(n + 1, dtype=np.uint32)
        rp[1:] = np.cumsum(counts)
        perm = rng.permutation(m)[:n]
        rows = []
        for i in range(n):
            others = rng.choice(m - 1, size=int(counts[i]) - 1, replace=False)
            rows.append(np.sort(np.concatenate([[perm[i]], others + (others >= perm[i])])))
        c = np.concatenate(rows).astype(np.uint32)
        v = rng.integers(1, 500, size=int(rp[-1])).astype(np.float64)
        eps = 0.5
    else:
        n, m, k = (70_000, 200_000, 16) if shape != "k64" else (20_000, 60_000, 64)
        rp, c, v = sla.generators.kregular_host(n, m, k, seed=5)
        if shape == "regular_f64":
            v = v + 0.125
        eps = None
    ref = oracle.jacobi_model("khosla", n, m, rp, c, v, eps=eps, khosla_scaling=False)
    solver, _ = sla.KhoslaSolver.new(n, m, len(c))
    solver.set_option("mesh_tail", tail)
    solver.load_csr(n, m, rp, c, v.copy())
    shard = MeshShard(solver, 0, 1, [0, n])
    shard.connect_pointers([shard.block])
    lo, hi, first = shard.local_value_range()
    for rep in range(3):
        shard.begin(False, eps, lo, hi, first)
        shard.solve()
        part = shard.finish()
        res = mesh_assemble([part], [shard.owned()], m)
        assert np.array_equal(res["p2o"], ref["p2o"]) and np.array_equal(res["o2p"], ref["o2p"])
        assert np.array_equal(res["prices"], ref["prices"])
        for key in ("num_unassigned", "bids", "bid_arcs", "dropped", "rounds"):
            assert res["stats"][key] == ref["stats"][key], (key, rep)
        assert shard.objective() == oracle_objective(rp, c, v, ref["p2o"])
    if shape == "regular_u16":
        assert solver.scan_value_bytes() == 2


def oracle_objective(rp, c, v, p2o):
    """Sum of the chosen arcs' values (all persons assigned, one arc per chosen column)."""
    n = len(rp) - 1
    counts = np.diff(rp.astype(np.int64))
    hit = np.asarray(c, dtype=np.int64) == np.repeat(p2o.astype(np.int64), counts)
    assert hit.sum() == n
    return float(np.asarray(v)[hit].sum())
